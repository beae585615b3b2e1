// tv5_internal.h — shared declarations of the libtv5 translation units (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "../../include/tv5.h"
#include "score.cuh"

#ifndef TV5_HYP_PER_THREAD
#define TV5_HYP_PER_THREAD 4
#endif
#ifndef TV5_SCORE_MINB
#define TV5_SCORE_MINB 2     // resident scoring CTAs per SM
#endif
#ifndef TV5_SCORE_UNROLL
#define TV5_SCORE_UNROLL 2   // point pairs per loop trip
#endif



namespace tv5 {

constexpr int kScoreThreads = 256;   // threads per scoring CTA
constexpr int kHypPerThread = TV5_HYP_PER_THREAD;  // hypotheses held in registers per thread
constexpr int kHypChunk = kScoreThreads * kHypPerThread;
constexpr int kScoreUnroll = TV5_SCORE_UNROLL;
#ifndef TV5_TILE_PAIRS
#define TV5_TILE_PAIRS 512
#endif
constexpr int kMaxTilePairs = TV5_TILE_PAIRS;   // point pairs staged in shared memory per tile (24 KB)
constexpr int kExactChunk = 2048;    // points per work item of the float64 scorer
constexpr int kHostChunks = 8;       // pipeline depth of the host-buffer entry point
constexpr int kPipeChunks = 8;       // chunks of pairs of one submission (solver / scorer overlap)
constexpr int kPipeMinPairs = 16;    // ... each at least this many pairs
constexpr int kProfPerChunk = 8;     // profiling events per chunk
constexpr int kEarlyMaxStages = 6;   // early exit: at most this many scoring stages

// Per image pair: geometry of the job (written by the host) ...
struct PairDesc {
  const double* x1;       // first point of the pair, [n,2]
  const double* x2;
  const int32_t* sets;    // [H,5], indices local to the pair
  double* E_out;          // [9]
  double* P_out;          // [12] or null
  tv5_result* result;
  uint8_t* mask_out;      // [n_full] or null
  int64_t pp_off;         // first PointPair32 of the pair in the workspace
  int32_t n;              // correspondences
  int32_t n_pre, n_full;
  int32_t pad;
};

// ... and its device-side state (zeroed at the start of every submission).
struct PairState {
  int32_t M;              // hypotheses produced by the solver (atomic)
  int32_t nonfinite;      // some coordinate is NaN/Inf
  uint32_t r1_bits, r2_bits;  // float bits of max ||(x1,1)||^2, ||(x2,1)||^2 (atomicMax)
  int32_t fast;           // 1: float32 guard-band scorer, 0: float64 scorer on every hypothesis
  int32_t tile_start;     // first scoring tile of this pair
  int32_t n_hc, n_pc;     // tiles = hypothesis chunks x point chunks
  int32_t n_cand;         // hypotheses to re-score exactly (atomic)
  int32_t exact_from;     // first candidate not yet re-scored
  int32_t n_entries;      // isolated real roots of the pair (split solver, atomic)
  int32_t M_total;        // hypotheses produced by the solver (M shrinks when early exit prunes)
  int32_t pp_lo, pp_hi;   // point pairs [pp_lo, pp_hi) scored by the current stage ...
  int32_t staged;         // ... when != 0; otherwise the whole pair
  int32_t pad2;
  double s_scale;         // sqrt(1-c)/thr folded into the float32 point / hypothesis records
  BandConst band;
};

struct Control {           // one per context, zeroed per submission
  int32_t n_tiles;
  int32_t tile_counter;
  int32_t pad[2];
};

struct Workspace {
  PairDesc* desc = nullptr;  size_t desc_cap = 0;     // [B]
  PairState* state = nullptr;                         // [B]
  Control* ctl = nullptr;
  PointPair32* pp = nullptr; size_t pp_cap = 0;       // [sum ceil(n/2)]
  double* E_list = nullptr;  size_t sets_cap = 0;     // [B*H,10,9]
  double* P_list = nullptr;                           // [B*H,10,12]
  int32_t* n_valid = nullptr;                         // [B*H]
  int32_t* n_roots = nullptr;                         // [B*H]
  double* rec = nullptr;                              // [B*H,96]  split solver: record per set
  void* entries = nullptr;                            // [B*H*10]  split solver: isolated roots (RootEntry)
  Hyp32* hyp = nullptr;      size_t hyp_cap = 0;      // [B*H*10]
  int32_t* hyp_id = nullptr;                          // [B*H*10]  set*16 + root
  uint32_t* notin = nullptr;                          // [B*H*10]
  uint32_t* out = nullptr;                            // [B*H*10]
  Hyp32* hyp2 = nullptr;                              // [B*H*10]  early exit: survivors (ping-pong)
  int32_t* hyp_id2 = nullptr;
  uint32_t* out2 = nullptr;
  int32_t* cand = nullptr;                            // [B*H*10]  hypothesis slots to re-score
  int32_t* cand_cnt = nullptr;                        // [B*H*10]  exact counts of the candidates
  int32_t* rng_sets = nullptr;                        // [B*H*5]   reference-RNG index tables of this submission
  // staging for the host-buffer entry point
  double* h2d_x = nullptr;   size_t h2d_cap = 0;      // [2 * sum n * 2]
  int32_t* h2d_sets = nullptr; size_t h2d_sets_cap = 0;
  double* out_E = nullptr;   size_t out_cap = 0;      // [B*9], [B*12], results
  double* out_P = nullptr;
  tv5_result* out_res = nullptr;
  // refinement (polish.cuh): job table, per-CTA partial sums, per-job barrier counters
  void* polish_jobs = nullptr; size_t polish_jobs_cap = 0;     // [jobs] PolishJob
  double* polish_partial = nullptr; size_t polish_partial_cap = 0;
  unsigned int* polish_barrier = nullptr;
  double* polish_x = nullptr; size_t polish_x_cap = 0;         // host-buffer entry point staging
  double* polish_E = nullptr;
  // flow -> correspondences (flow_points.cuh)
  void* flow_jobs = nullptr; size_t flow_jobs_cap = 0;         // [B] FlowJob
  double* flow_x = nullptr; size_t flow_x_cap = 0;             // x1 | x2 of tv5_pose_from_flow
  double* flow_EP = nullptr; size_t flow_EP_cap = 0;           // [B,9] | [B,12] float64 results
};

// single-pair launch sequence captured as a CUDA graph, one per shape
struct GraphKey {
  int n, iters, cheir, flags;
  double thr;
  bool operator==(const GraphKey& o) const {
    return n == o.n && iters == o.iters && cheir == o.cheir && flags == o.flags && thr == o.thr;
  }
};
struct GraphEntry { GraphKey key; cudaGraphExec_t exec; };
struct GraphSeen { GraphKey key; int count; };
struct GuardRec { void* user; void* base; size_t bytes; size_t payload; };   // guard mode: see ws_malloc_bytes

}  // namespace tv5

struct tv5_ctx {
  std::recursive_mutex mu;               // held by every entry point: host threads sharing a context are serialised
  int device = 0;
  int sm_count = 0;
  int last_cuda = 0;
  tv5::Workspace ws;
  float* rng_u = nullptr;               // uniform draws of the reference RNG, [rng_u_iters*5][512]
  int rng_u_iters = 0;
  cudaStream_t last_stream = nullptr;   // stream of the previous submission (cross-stream ordering)
  cudaEvent_t last_done = nullptr;
  bool has_last = false;
  bool force_exact = false;
  bool guard = false;                   // guard zones + poisoned payloads on every workspace buffer (testing aid)
  int poison = 0;
  std::vector<tv5::GuardRec> guards;
  bool profiling = false;
  cudaEvent_t ev[TV5_N_STAGES + 1] = {};
  double stage_ms[TV5_N_STAGES] = {};
  int64_t stage_launches[TV5_N_STAGES] = {};
  bool ev_pending = false;
  cudaStream_t copy_stream = nullptr;   // host-buffer entry point: H2D copies overlap compute
  cudaEvent_t chunk_ev[tv5::kHostChunks] = {};
  cudaEvent_t start_ev = nullptr;
  int polish_max_ctas = 0;              // co-resident CTAs of irls_polish on this device
  // solver / scorer overlap inside one submission
  bool use_graphs = true;               // single-pair submissions replay a captured CUDA graph
  std::vector<tv5::GraphEntry> graphs;
  std::vector<tv5::GraphSeen> graph_seen;   // shapes seen so far; captured on the third occurrence
  cudaStream_t cap_stream = nullptr;
  cudaStream_t cap_stream2 = nullptr;   // second branch while capturing (point preparation next to the solver front)
  cudaEvent_t cap_fork = nullptr, cap_join = nullptr;
  bool early_exit = false;              // staged scoring with exact hypothesis pruning (opt-in)
  int early_stages = 3;                 // ... stage boundaries as fractions of a pair's points
  float early_frac[tv5::kEarlyMaxStages + 1] = {0.0f, 0.30f, 0.52f, 1.0f};
  bool split_solver = true;             // three-kernel solver (solve5_split.cuh) instead of solve_sets
  bool overlap = false;
  bool profiling_serial = false;
  cudaStream_t front_stream = nullptr;  // prep + solve, least priority
  cudaStream_t back_stream = nullptr;   // scoring + selection, highest priority
  cudaEvent_t pipe_entry = nullptr, pipe_done = nullptr;
  cudaEvent_t pipe_solved[tv5::kPipeChunks] = {};
  std::vector<cudaEvent_t> prof_ev;
  int prof_chunks = 0;
};
