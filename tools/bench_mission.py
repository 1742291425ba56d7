"""BASELINE configs 1 and 2 end to end: whole missions on the device-backed stack, with the CPU oracle planner timed at
snapshots of the same mission.

  config 1  myopic_experiments.py: UCB reward, 10x10 Gaussian world (RBF l=1, var=100), Dubins fan (frontier 15,
            horizon 1.5, sample step 0.5, turning radius 0.05), noise 1e-4, 150 steps
  config 2  nonmyopic_experiments.py: cMCTS dpw, MES reward (max-value samples every step), horizon 5, 250 rollouts per
            step, noise 0.5, 250 steps (the reference's script runs 150)
Per step (Robot.planner / choose_trajectory, robot_library.py:155-340): predicted maximum over a 100 x 100 grid, action
choice, observations from Environment.sample_value along the chosen path, belief update.
"""
import contextlib, io, json, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

R = (0., 10., 0., 10.)


def grid100():
    x1, x2 = np.meshgrid(np.linspace(R[0], R[1], 100), np.linspace(R[2], R[3], 100))
    return np.vstack([x1.ravel(), x2.ravel()]).T


def run_mission(kind, T, noise, seed=0, cpu_snapshots=()):
    import torch
    from informative_path_planning_b200 import aq_library as aq, gpmodel_library as gp, mcts_library as mc
    from informative_path_planning_b200.envmodel_library import Environment
    from informative_path_planning_b200.paths_library import Dubins_Path_Generator
    from oracle import aq_oracle, mcts_oracle
    from oracle.gpmodel_oracle import OnlineGPModelOracle
    world = Environment(R, 20, 100.0, 1.0, noise=noise, seed=seed, dim=2)
    model = gp.OnlineGPModel(R, 1.0, 100.0, noise=noise)
    pg = Dubins_Path_Generator(15, 1.5, 0.05, 0.5, R)
    pose = (5.0, 5.0, 0.0)
    grid = grid100()
    np.random.seed(seed + 100); random.seed(seed + 200)
    xobs = np.array([[pose[0], pose[1]]])
    model.add_data(xobs, world.sample_value(xobs))                      # Robot.__init__ takes a first observation
    step_s, plan_s, cpu = [], [], []
    evals = 0
    t_all = time.perf_counter()
    for t in range(T):
        t0 = time.perf_counter()
        mu, _ = model.predict_value(grid)                               # predict_max (robot_library.py:222-254)
        t1 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
            if kind == 'myopic':
                best = mc.choose_myopic(model, pg, pose, 'mean', t)
                if best is None:
                    break
                path, dense = best[0], best[1]
                evals += len(best[3])
            else:
                planner = mc.cMCTS(250, model, pose, 5, pg, aq.mves, 'mes', t, tree_type='dpw')
                out = planner.choose_trajectory(t=t)
                path, dense = out[0], out[1]
                evals += planner.tree.reward_evals
        t2 = time.perf_counter()
        cpu_dt = 0.0
        if t in cpu_snapshots:                                          # the CPU oracle planner on the same belief
            o = OnlineGPModelOracle(R, 1.0, 100.0, noise=noise)
            o.add_data(model.xvals, model.zvals)
            st_np, st_py = np.random.get_state(), random.getstate()
            c0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
                o.predict_value(grid)
                if kind == 'myopic':
                    paths, _ = pg.get_path_set(pose)
                    for k in paths:
                        aq_oracle.mean_UCB(t, paths[k], o)
                else:
                    mcts_oracle.cMCTSOracle(250, o, pose, 5, pg, aq_oracle.mves, 'mes', t, tree_type='dpw').choose_trajectory(t=t)
            cpu_dt = time.perf_counter() - c0
            cpu.append({'t': t, 'n_obs': int(model.n_obs), 'cpu_step_s': cpu_dt, 'device_step_s': t2 - t0})
            np.random.set_state(st_np); random.setstate(st_py)
        xobs = np.array(path)[:, :2]
        try:
            model.add_data(xobs, world.sample_value(xobs))
        except np.linalg.LinAlgError:
            break
        pose = dense[-1]
        torch.cuda.synchronize()
        step_s.append(time.perf_counter() - t0 - cpu_dt)
        plan_s.append(t2 - t1)
    total = time.perf_counter() - t_all - sum(c['cpu_step_s'] for c in cpu)
    return {'kind': kind, 'steps_done': len(step_s), 'steps_asked': T, 'noise': noise, 'n_obs_final': int(model.n_obs),
            'mission_s': total, 'steps_per_s': len(step_s) / max(total, 1e-9), 'median_step_ms': 1e3 * float(np.median(step_s)),
            'last10_step_ms': 1e3 * float(np.mean(step_s[-10:])), 'median_plan_ms': 1e3 * float(np.median(plan_s)),
            'reward_evals': evals, 'reward_evals_per_s': evals / max(sum(plan_s), 1e-9), 'cpu_snapshots': cpu}


if __name__ == '__main__':
    out = []
    run_mission('myopic', 5, 1e-4)                                      # warm-up (module load, allocator)
    out.append(run_mission('myopic', 150, 1e-4, cpu_snapshots=(10, 75, 149)))
    print(json.dumps(out[-1]))
    out.append(run_mission('myopic', 150, 0.5, cpu_snapshots=(149,)))
    print(json.dumps(out[-1]))
    out.append(run_mission('nonmyopic', 250, 0.5, cpu_snapshots=(5, 60, 249)))
    print(json.dumps(out[-1]))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(out, open('gpurun_out/mission_bench.json', 'w'), indent=1)
