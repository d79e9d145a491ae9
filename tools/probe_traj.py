"""trajectory_config3 leg alone (k-means++ start, default vs bounded + incremental); SKM_PRUNE=0/1 via the environment."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rig = bench.Rig()
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "config3"]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n, p, K, m = cfg["n"], cfg["p"], cfg["K"], cfg["m"]
ds, views, mu, start = bench.gen_dataset(rig.ctx, rig.dev, n, p, m, K, col0=0, kind="mixture")
t = bench.trajectory_leg(rig, ds, cfg, n, max_iter=iters)
print(json.dumps({"SKM_PRUNE": os.environ.get("SKM_PRUNE"), "speedup": t["speedup"], "default_total": t["default"]["total_ms"],
                  "modes_total": t["bounded_incremental"]["total_ms"],
                  "default_ms": [round(x, 2) for x in t["default"]["ms_per_iteration"]],
                  "modes_ms": [round(x, 2) for x in t["bounded_incremental"]["ms_per_iteration"]],
                  "identical": t["iterations_with_identical_assignments"]}))
