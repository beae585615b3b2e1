import sys, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench
r = bench.config5_report()
print(json.dumps({k: ({kk: vv for kk, vv in v.items()} if isinstance(v, dict) else "...") for k, v in r.items() if k != "workload"}, indent=0)[:3500])
