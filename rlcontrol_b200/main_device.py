"""``main.py`` of the reference for the runs the device-resident loop covers.

Same command line (``--env_json --agent_json --indices START STEP STOP --save_dir``, main.py:31-42), same INDEX ->
setting / run / seed decoding (main.py:113-141, utils/main_utils.py:90-98), same pickle: ``data["experiment"]`` metadata
and ``data["experiment_data"][setting] = {"agent_params", "runs": [run_data, ...]}`` written to
``<save_dir>/<env>_<agent>results/data_<START>_<STEP>_<STOP>.pkl`` (main.py:80-95,188-203).  Instead of one OS process
per INDEX (main_concurrent.py:64-81) the runs of an invocation are packed ``--runs_per_gpu`` at a time onto this
process's GPU, each with its own stream and captured graphs (replicas only, no communication); under ``torchrun`` rank
r takes every WORLD_SIZE-th INDEX and writes its own file, named after the (start, step, stop) triple it actually ran.

    python -m rlcontrol_b200.main_device --env_json jsonfiles/environment/Pendulum-v0.json \\
        --agent_json jsonfiles/agent/reverse_kl.json --indices 0 1 64 --save_dir ./results
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
from collections import OrderedDict

from . import sweep
from .device_loop import DeviceExperiment, EnvSpec, run_interleaved

AGENTS = ("ReverseKL", "ForwardKL")


def run_indices(env_json: dict, agent_json: dict, indices, save_dir: str, env_name: str, agent_name: str,
                runs_per_gpu: int = 8, device: int | None = None, extra: dict | None = None, verbose: bool = True):
    """Run INDEX = range(*indices) and write the reference's pickle; returns (path, data)."""
    import torch

    from . import kl_networks
    from .engine import Engine
    if agent_json["agent"] not in AGENTS:
        raise NotImplementedError("device-resident runs cover %s; got %r" % (AGENTS, agent_json["agent"]))
    cls = kl_networks.ReverseKLNetwork if agent_json["agent"] == "ReverseKL" else kl_networks.ForwardKLNetwork
    spec = EnvSpec(env_json)
    data = {"experiment": {"environment": {}, "agent": {}}, "experiment_data": {}}
    data["experiment"]["agent"]["agent_name"] = agent_json["agent"]
    data["experiment"]["agent"]["parameters"] = dict(agent_json["sweeps"])
    e = data["experiment"]["environment"]
    e["env_name"] = env_json["environment"]
    e["total_timesteps"] = env_json["TotalMilSteps"] * 1000000
    e["steps_per_episode"] = env_json["EpisodeSteps"]
    e["eval_interval_timesteps"] = env_json["EvalIntervalMilSteps"] * 1000000
    e["eval_episodes"] = env_json["EvalEpisodes"]
    start, step, stop = (int(x) for x in indices)
    todo = list(range(start, stop, step))
    for g0 in range(0, len(todo), max(1, int(runs_per_gpu))):
        group, exps = todo[g0:g0 + runs_per_gpu], []
        for index in group:
            params, total = sweep.sweep_setting(agent_json["sweeps"], index)
            setting, run, seed = sweep.index_to_run(index, total)
            cfg = sweep.make_config(spec.env_params(), dict(params), dict(random_seed=seed, write_log=False, write_plot=False),
                                    engine=Engine(device), **(extra or {}))
            torch.manual_seed(seed)            # the reference leaves torch unseeded; a run here is reproducible
            exps.append((index, setting, dict(params), DeviceExperiment(cls(None, None, cfg), env_json, cfg)))
        run_interleaved([x[3] for x in exps])
        for index, setting, params, exp in exps:
            slot = data["experiment_data"].setdefault(setting, {"agent_params": params, "runs": []})
            slot["runs"].append(exp.run_data(env_json))
            if verbose:
                r = exp.eval_rewards_per_episode[-1]
                print("index %d (setting %d, seed %d): %d episodes, last evaluation mean return %.1f, %.1f s" %
                      (index, setting, exp.seed, exp.train_episodes, sum(r) / len(r), exp.wall), flush=True)
    out_dir = save_dir + "/" + env_name + "_" + agent_name + "results/"
    os.makedirs(out_dir, exist_ok=True)
    path = out_dir + f"data_{start}_{step}_{stop}.pkl"
    with open(path, "wb") as f:
        pickle.dump(data, f)
    return path, data


def rank_indices(indices, rank: int, world: int):
    """(start, step, stop) of the INDEX values rank ``rank`` of ``world`` runs: every world-th one, no overlap."""
    start, step, stop = (int(x) for x in indices)
    return start + rank * step, step * world, stop


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--env_json", type=str, required=True)
    p.add_argument("--agent_json", type=str, required=True)
    p.add_argument("--indices", type=int, nargs=3, required=True)
    p.add_argument("--save_dir", default="./results")
    p.add_argument("--runs_per_gpu", type=int, default=8)
    args = p.parse_args(argv)
    # main.py:50-51 names the result directory with str.rstrip(".json") (a character-set strip); kept
    env_name = os.path.basename(args.env_json).rstrip(".json")
    agent_name = os.path.basename(args.agent_json).rstrip(".json")
    with open(args.env_json) as f:
        env_json = json.load(f, object_pairs_hook=OrderedDict)
    with open(args.agent_json) as f:
        agent_json = json.load(f, object_pairs_hook=OrderedDict)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    mine = rank_indices(args.indices, rank, world)             # every WORLD_SIZE-th INDEX; no communication
    import torch
    torch.cuda.set_device(local)
    path, _ = run_indices(env_json, agent_json, mine, args.save_dir, env_name, agent_name, args.runs_per_gpu, device=local)
    print("rank %d wrote %s" % (rank, path), flush=True)


if __name__ == "__main__":
    main()
