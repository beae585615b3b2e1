// Dev tool: per-phase cycle counts of the per-thread solver (one warp per SM-ish, like a single pair).
#define TV5_SOLVE_PROFILE 1
#include <cstdio>
#include <vector>
#include <random>
#include "../deep-sfm-revisited_b200/csrc/solve5_coop.cuh"
using namespace tv5;
__global__ void __launch_bounds__(32, MINB) k(const double* x1, const double* x2, const int* sets, int H, double* E, double* P, int* nv) {
  __shared__ double sB[kCoopBasisDoubles][kCoopStride]; __shared__ double sR[kCoopRowsDoubles][kCoopStride]; __shared__ double sQ[kCoopPointDoubles][kCoopStride]; __shared__ int sOk[32];
  int h = blockIdx.x * blockDim.x + threadIdx.x; bool valid = h < H; if (!valid) h = H - 1;
  double q[5][2], qp[5][2];
  for (int i = 0; i < 5; ++i) { int idx = sets[5*h+i]; q[i][0]=x1[2*idx]; q[i][1]=x1[2*idx+1]; qp[i][0]=x2[2*idx]; qp[i][1]=x2[2*idx+1]; }
  int nr; int v = solve_minimal_set_coop(valid, q, qp, true, E + 90*(size_t)h, P + 120*(size_t)h, &nr, sB, sR, sQ, sOk); if (valid) nv[h] = v;
}
int main(int argc, char** argv) {
  int H = argc > 1 ? atoi(argv[1]) : 4096, N = 10000;
  std::mt19937 rng(1); std::uniform_real_distribution<double> U(-0.8, 0.8), D(5, 80); std::normal_distribution<double> G(0, 7e-5);
  std::vector<double> x1(2*N), x2(2*N);
  for (int i = 0; i < N; ++i) { double u=U(rng), v=0.3*U(rng), z=D(rng); double X=u*z, Y=v*z, Z=z-0.8; x1[2*i]=u; x1[2*i+1]=v; x2[2*i]=(X+0.03)/Z+G(rng); x2[2*i+1]=(Y-0.01)/Z+G(rng); }
  std::vector<int> sets(5*H); for (auto& s : sets) s = rng() % N;
  double fill[4][9]; double ran = 3.18730379; for (int i=0;i<4;++i) for (int j=0;j<9;++j){ ran*=3.18730379; ran=2.0*(ran-floor(ran))-1.0; fill[i][j]=ran; }
  cudaMemcpyToSymbol(c_completion, fill, sizeof(fill));
  double *dx1,*dx2,*E,*P; int *ds,*nv;
  cudaMalloc(&dx1,16*N); cudaMalloc(&dx2,16*N); cudaMalloc(&ds,20*H); cudaMalloc(&E,720*(size_t)H); cudaMalloc(&P,960*(size_t)H); cudaMalloc(&nv,4*H);
  cudaMemcpy(dx1,x1.data(),16*N,cudaMemcpyHostToDevice); cudaMemcpy(dx2,x2.data(),16*N,cudaMemcpyHostToDevice); cudaMemcpy(ds,sets.data(),20*H,cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; ++rep) {
    unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_solve_prof, z, sizeof(z));
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0);
    k<<<(H+31)/32, 32>>>(dx1,dx2,ds,H,E,P,nv);
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1);
    unsigned long long p[16]; cudaMemcpyFromSymbol(p, g_solve_prof, sizeof(p));
    const char* names[8] = {"nullspace","coop c+elim","(unused)","determinant","roots","E+cheirality","  roots:build","  roots:isolate"};
    double tot = 0; for (int i=0;i<6;++i) tot += p[i];
    printf("H=%d  %.3f ms  (%s)\n", H, ms, cudaGetErrorString(cudaGetLastError()));
    for (int i=0;i<8;++i) printf("  %-14s %8.0f cycles/warp  %5.1f%%\n", names[i], (double)p[i]/((H+31)/32), 100.0*p[i]/tot);
  }
  { unsigned long long p[16]; cudaMemcpyFromSymbol(p, g_solve_prof, sizeof(p)); printf("per set: isolate trips %.2f, sturm evals %.2f, newton iters %.2f, roots refined %.2f\n", (double)p[8]/H, (double)p[9]/H, (double)p[10]/H, (double)p[11]/H); }
  std::vector<int> hnv(H); cudaMemcpy(hnv.data(), nv, 4*H, cudaMemcpyDeviceToHost); long s=0; for (int v: hnv) s+=v; printf("mean n_valid %.3f\n", (double)s/H);
  return 0;
}
