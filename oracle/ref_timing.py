"""Times the LIVE Python reference in the build container (no GPU, needs /root/reference):
  (1) env-only: Scene.run_game('intelligent') driven by an A*-guided controller with the
      limited-visual StateManager encode in both callbacks;
  (2) AG-DQN training: the reference's MADQNAgentController.model_run (train_NN) on CPU.
Writes profiles/r01_reference_python_timing.json.  TEST/REPORT INFRASTRUCTURE ONLY."""
import contextlib
import io
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle.oracle import README_9x9  # noqa: E402


def env_only(seconds=20.0):
    R = rh.load()
    turns, t0 = 0, time.perf_counter()
    seed = 0
    while time.perf_counter() - t0 < seconds:
        scene = rh.make_scene(4, seed=seed, layout_list=README_9x9.tolist())
        ctl = rh.RecordingController(scene, random.Random(seed), p_guided=1.0, n_actions=4)
        orig_store = ctl.store_info

        def store(all_info, reward, is_end, name, ctl=ctl, orig=orig_store):
            ctl.sm.create_state(all_info, name, obs_clip=True)  # next_obs encode, like MADQN Controller.py:139
            orig(all_info, reward, is_end, name)
        ctl.store_info = store
        rh.run_intelligent(scene, ctl)
        turns += len(ctl.turns)
        seed += 1
    dt = time.perf_counter() - t0
    return {"agv_steps": turns, "seconds": dt, "agv_steps_per_s": turns / dt, "episodes": seed,
            "what": "9x9 README layout, 4 AGVs, intelligent mode, A* guidance + 2 clip encodes per turn, 1 core"}


def agdqn_training(seconds=60.0):
    R = rh.load()
    sys.path.insert(0, os.path.join(rh.REF_ROOT, "src", "algorithm"))
    import importlib
    with contextlib.redirect_stdout(io.StringIO()):
        C = importlib.import_module("src.algorithm.MADQN_structure.Controller")
    import torch
    torch.manual_seed(0)
    random.seed(0)
    scene = rh.make_scene(1, seed=0, layout_list=README_9x9.tolist())
    orig_run = scene.run_game
    scene.run_game = lambda control_pattern="manual", smart_controller=None, render=True: orig_run(
        control_pattern=control_pattern, smart_controller=smart_controller, render=False)
    with contextlib.redirect_stdout(io.StringIO()):
        ctl = C.MADQNAgentController(scene, 9, 9, len(scene.layout.storage_station_list), control_mode="train_NN",
                                     state_number=3)
    ctl.storage_path = "/tmp"
    steps, episodes, t0 = 0, 0, time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        while time.perf_counter() - t0 < seconds:
            ctl.self_init()
            scene.init()
            try:
                scene.run_game(control_pattern="intelligent", smart_controller=ctl, render=True)
            except IndexError:
                pass  # STOP (4) indexes the 4-wide Q row, SURVEY 0.5
            steps += ctl.action_length_record
            episodes += 1
    dt = time.perf_counter() - t0
    return {"agv_steps": steps, "seconds": dt, "agv_steps_per_s": steps / dt, "episodes": episodes,
            "updates": ctl.agent.learn_step_counter, "torch_threads": torch.get_num_threads(),
            "what": "reference MADQNAgentController train_NN, 9x9 README layout, 1 AGV, CPU torch; one optimiser "
                    "step (batch 32) per stored transition after 100 warm-up transitions"}


if __name__ == "__main__":
    out = {"env_only": env_only(), "agdqn_training": agdqn_training(),
           "host": "build container, %d cores" % (os.cpu_count() or 0)}
    path = os.path.join(ROOT, "profiles", "r01_reference_python_timing.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
