"""Golden decision traces of the reference's block-size scheduler (benchmark_dynamic_schedule.py:54-257, "ewma" mode).

Run in the build container only (needs /root/reference):  python tests/golden/make_sched_golden.py
The reference class is taken from its source file as is (the module around it imports datasets / tokenizers that a
scheduler trace does not need, so only the class statement is executed). Writes tests/golden/scheduler_traces.json;
tests/test_schedule_cpu.py replays the same (tau, cycle time) streams through dflash_b200.schedule.EwmaBlockScheduler.
"""
import ast
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/benchmark_dynamic_schedule.py"

CASES = [
    dict(candidates=[8, 16, 32], warmup_cycles=6, ewma_alpha=0.3, switch_margin=0.05, required_streak=2,
         cooldown_cycles=3, probe_interval=7, low_accept_threshold=0.15, low_accept_streak=3, seed=1, n=120),
    dict(candidates=[4, 8, 12, 16], warmup_cycles=0, ewma_alpha=0.5, switch_margin=0.0, required_streak=1,
         cooldown_cycles=0, probe_interval=3, low_accept_threshold=0.3, low_accept_streak=2, seed=2, n=150),
    dict(candidates=[16], warmup_cycles=2, ewma_alpha=1.0, switch_margin=0.1, required_streak=2, cooldown_cycles=1,
         probe_interval=2, low_accept_threshold=0.5, low_accept_streak=1, seed=3, n=20),
    dict(candidates=[8, 16], warmup_cycles=4, ewma_alpha=0.2, switch_margin=0.2, required_streak=3, cooldown_cycles=5,
         probe_interval=0, low_accept_threshold=0.0, low_accept_streak=3, seed=4, n=80),
]


def workload(seed, n):
    """(tau(b), cycle_s(b)) generator: acceptance saturates with the block size, cycle time grows with it; the
    regime drifts over time so that the best block size changes."""
    rng = random.Random(seed)
    for i in range(n):
        cap = 3.0 + 10.0 * (0.5 + 0.5 * (1 if (i // 40) % 2 == 0 else -1) * 0.8)
        yield i, rng, cap


def trace(cls, case, **extra):
    kw = {k: v for k, v in case.items() if k not in ("seed", "n")}
    s = cls(**kw, **extra)
    sel, cur = [], []
    for i, rng, cap in workload(case["seed"], case["n"]):
        b = s.select(i)
        tau = max(1, min(b, int(rng.expovariate(1.0 / cap)) + 1))
        eff = b if i % 17 != 16 else 1  # an occasional clamped tail block
        cyc = 0.010 + 0.0004 * b + rng.random() * 0.001
        s.update(tau=tau, cycle_s=cyc, effective_bs=eff, cycle_idx=i)
        sel.append(b)
        cur.append(s.current)
    return dict(select=sel, current=cur)


def main():
    import numpy as np
    tree = ast.parse(open(SRC).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "EWMAPerformanceScheduler")
    ns = {"np": np}
    exec(compile(ast.Module([node], []), SRC, "exec"), ns)
    ref = ns["EWMAPerformanceScheduler"]
    out = []
    for case in CASES:
        t = trace(ref, case, scheduler_mode="ewma", adl_rho=0.5, adl_delta=1.0, adl_k_min=2, adl_k_max=32,
                  adl_neighborhood=1)
        out.append(dict(case=case, **t))
    json.dump(out, open(os.path.join(HERE, "scheduler_traces.json"), "w"))
    print("wrote", len(out), "traces;", [len(set(t["current"])) for t in out], "distinct current sizes per case")


if __name__ == "__main__":
    main()
