"""Tree search with the reference's heuristic rewards (naive / naive_value: run_sim_all.sh:17): rollouts/s of one planning
step (dpw, horizon 5, 250 rollouts, Dubins fan) -- native / Python tree without belief updates / the reference's
per-level sequence on the device / CPU oracle planner."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_mcts
out = []
for f_rew in ('naive', 'naive_value'):
    for n_obs in (100, 900):
        rows = []
        for kind, fuse, native in (('gpu', True, True), ('gpu', True, False), ('gpu', False, False), ('cpu', True, False)):
            if kind == 'gpu':
                bench_mcts.run(kind, n_obs, f_rew, budget=20, fuse=fuse, native=native)
            reps = 3 if kind == 'gpu' else 1
            r = min((bench_mcts.run(kind, n_obs, f_rew, fuse=fuse, native=native) for _ in range(reps)), key=lambda x: x['seconds'])
            rows.append(r)
            print(json.dumps({k: r[k] for k in ('kind', 'n_obs', 'f_rew', 'seconds', 'rollouts_per_s', 'root_visits')}))
        print('same root visit counts (device variants):', rows[0]['root_visits'] == rows[1]['root_visits'] == rows[2]['root_visits'])
        out += rows
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/mcts_naive_bench.json', 'w'), indent=1)
