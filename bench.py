#!/usr/bin/env python3
"""bench.py -- SBMF Gibbs sweep throughput on B200 (the driver's bench contract; see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W            our CUDA path (through the C ABI, libsbmf_cuda.so)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU program (oracle/_ref), host cores

Workload (BASELINE.json metric: "Gibbs sweeps/s and |Omega|*K factor-updates/s, Netflix-100M K=100"): a Netflix-shaped
synthetic rating matrix, 480,189 x 17,770, ~100.48M TRAIN ratings (+10% held out for the test RMSE), K=100, generated
on the device by sbmf_cuda_synth_generate.  One step = one full Gibbs sweep = one body of gibbs_sbpmf2.cpp:335-637
(residual rebuild, all hyper-parameters, user phase, item phase, test prediction + RMSE).
value  = |Omega_train| * K * sweeps/s with everything resident in HBM, timed by CUDA events on the library's stream.
e2e    = the same metric for the whole job through the C ABI from HOST buffers: set_train (H2D of the COO triples + device
         CSR/CSC build) + set_test + init_factors + K x (sweep + D2H read of the RMSE) + D2H of the predictions.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200"))
# rank 0 prints exactly ONE JSON line on stdout: anything native libraries print (e.g. NCCL's version banner) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

WORKLOADS = {
    # name: (users, items, train ratings, K)
    "netflix_k100": (480189, 17770, 100480507, 100),
    "netflix_k50": (480189, 17770, 100480507, 50),
    "netflix_k200": (480189, 17770, 100480507, 200),
    "ml10m_k100": (71567, 10681, 10000000, 100),
    "ml20m_k200": (138493, 26744, 20000000, 200),
    "ml1m_k50": (6040, 3706, 1000209, 50),
    # BASELINE.json configs[4]; the 1e13-pair grid is sampled sparsely on host threads (csrc/synth_host.cpp).  Fits one B200
    # (~70 GB at the peak of the layout build); use --steps 3 or so: a sweep takes about a second on one GPU
    "scaled_1b_k200": (10000000, 1000000, 900000000, 200),
}


def generate(sbmf, workload, device):
    I, J, NTRAIN, _ = WORKLOADS[workload]
    n = int(round(NTRAIN / (1 - TEST_FRAC)))
    if float(I) * float(J) > 1e11:
        return sbmf.synth_generate_host(I, J, n, test_frac=TEST_FRAC, seed=SEED)
    return sbmf.synth_generate(I, J, n, test_frac=TEST_FRAC, seed=SEED, device=device)
TEST_FRAC = 0.1
SEED = 20151001 + 3
METRIC = "gibbs_factor_updates_per_s"
UNIT = "factor-updates/s"
ALG_BYTES_PER_FU = 24.0          # 12 B per (rating, dimension) visit and half-step, two half-steps


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev=0):
        self.dev, self.proc, self.lines = dev, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own gibbs_sbpmf2.cpp (oracle/_ref, built from /root/reference by
# `make -C oracle ref`), else the oracle port.  A bounded sample of the workload: the first users of the same matrix.
def write_triples(path, u, i, r):
    with open(path, "w") as f:
        np.savetxt(f, np.column_stack([u.astype(np.int64), i.astype(np.int64), r.astype(np.float64)]), fmt="%d\t%d\t%g")


def take_sample(d, n_target):
    """user-sorted prefix of the train set with ~n_target ratings, and those users' test ratings"""
    tu = d["train_user"]
    n = min(n_target, tu.size)
    last_user = int(tu[n - 1])
    n = int(np.searchsorted(tu, last_user, side="left")) or n      # whole users only
    umax = int(tu[n - 1])
    nt = int(np.searchsorted(d["test_user"], umax, side="right"))
    return {"train_user": tu[:n], "train_item": d["train_item"][:n], "train_rating": d["train_rating"][:n],
            "test_user": d["test_user"][:nt], "test_item": d["test_item"][:nt], "test_rating": d["test_rating"][:nt],
            "num_users": umax + 1, "num_items": d["num_items"]}


def run_reference_binary(sample, K, threads, pairs=1, timeout=600):
    """per-sweep seconds of the unmodified reference = (wall(T=Tb) - wall(T=1)) / (Tb - 1): load time cancels."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from run_ref import ref_binary, run_ref
    Tb = 3 if ref_binary(K, 3) else 2
    b1, b3 = ref_binary(K, 1), ref_binary(K, Tb)
    if not b1 or not b3:
        return None
    with tempfile.TemporaryDirectory(prefix="sbmf_bench_") as tmp:
        tr, te = os.path.join(tmp, "train"), os.path.join(tmp, "test")
        # the reference sizes its arrays by max id over train U test: make the item id space explicit with one test row
        write_triples(tr, sample["train_user"], sample["train_item"], sample["train_rating"])
        tu = np.append(sample["test_user"], sample["num_users"] - 1)
        ti = np.append(sample["test_item"], sample["num_items"] - 1)
        trr = np.append(sample["test_rating"], 3.0)
        write_triples(te, tu, ti, trr)
        per = []
        rm = None
        for _ in range(pairs):
            r1 = run_ref(b1, tr, te, threads=threads, timeout=timeout)
            r3 = run_ref(b3, tr, te, threads=threads, timeout=timeout)
            if len(r1["rmse"]) != 1 or len(r3["rmse"]) != Tb:
                return None
            per.append((r3["wall_s"] - r1["wall_s"]) / (Tb - 1))
            rm = r3["rmse"]
    return {"s_per_sweep": float(np.median(per)), "sweeps_timed": pairs * (Tb - 1), "rmse": rm}


def run_oracle_port(sample, K, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as orc
    o = orc.Oracle(sample["train_user"], sample["train_item"], sample["train_rating"], sample["test_user"], sample["test_item"],
                   sample["test_rating"], sample["num_users"], sample["num_items"], K, noise=orc.NOISE_RAND)
    o.init_factors()
    o.sweep(warmup)
    t0 = time.perf_counter()
    r, _ = o.sweep(steps)
    dt = time.perf_counter() - t0
    return {"s_per_sweep": dt / steps, "sweeps_timed": steps, "rmse": [float(x) for x in r]}


def cpu_reference(sample, K, pairs=1, port_steps=2):
    """Returns the cpu_baseline object: best of 1 thread and all host threads of the reference's OpenMP build."""
    n = int(sample["train_user"].size)
    ncpu = os.cpu_count() or 1
    desc = f"first {sample['num_users']} users of the workload ({n} train ratings, {sample['num_items']} items, K={K})"
    out = None
    try:
        r1 = run_reference_binary(sample, K, 1, pairs)
    except Exception as e:   # noqa: BLE001 - the reference binary may be absent or time out; fall back to the port
        r1 = None
        desc += f"; reference binary failed: {type(e).__name__}"
    if r1:
        runs = {1: r1}
        if ncpu > 1:
            try:
                rn = run_reference_binary(sample, K, ncpu, 1, timeout=max(120.0, 40 * r1["s_per_sweep"] + 60))
                if rn:
                    runs[ncpu] = rn
            except Exception:   # noqa: BLE001 - glibc rand() lock contention can make the OpenMP run pathologically slow
                desc += f"; OMP_NUM_THREADS={ncpu} run timed out (rand() lock contention, SURVEY.md 0.8)"
        best = min(runs, key=lambda t: runs[t]["s_per_sweep"])
        out = {"value": n * K / runs[best]["s_per_sweep"], "unit": UNIT, "cores": best, "kind": "reference",
               "sample": desc + f"; unmodified gibbs_sbpmf2.cpp (g++ -O3 -fopenmp), per-sweep time = (wall(T=3)-wall(T=1))/2",
               "by_threads": {str(t): n * K / v["s_per_sweep"] for t, v in runs.items()}, "host_cpus": ncpu,
               "sweeps_timed": runs[best]["sweeps_timed"], "s_per_sweep": runs[best]["s_per_sweep"]}
    else:
        r = run_oracle_port(sample, K, port_steps, 1)
        out = {"value": n * K / r["s_per_sweep"], "unit": UNIT, "cores": 1, "kind": "port",
               "sample": desc + "; oracle/sbmf_oracle.c (scalar C restatement of gibbs_sbpmf2.cpp), oracle/_ref binaries absent",
               "host_cpus": ncpu, "sweeps_timed": r["sweeps_timed"], "s_per_sweep": r["s_per_sweep"]}
    return out


# ---------------------------------------------------------------------------------------------------------------------
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netflix_k100", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=2000000, help="train ratings in the CPU baseline's sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    rank, local_rank, world = dist_env()
    if world != a.gpus and world != 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    I, J, NTRAIN, K = WORKLOADS[a.workload]
    W = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    cfg = {"workload": f"{a.workload}: synthetic Netflix/MovieLens-shaped Zipf rating matrix {I}x{J}, ~{NTRAIN} train ratings (+{int(TEST_FRAC * 100)}% test), K={K}",
           "users": I, "items": J, "K": K, "l2_policy": "working set per sweep (>2 GB) far exceeds the 126 MB L2; no flush needed",
           "rebuild_every": 1, "residual_mode": "0: per-sweep residual rebuild of gibbs_sbpmf2.cpp:342-359 fused into the user phase",
           "sample_mode": "reference (x = mu + (1/lambda) z)"}

    import sbmf
    if os.environ.get("SBMF_EMULATED") or hasattr(sbmf.load_library(), "sbmf_simt_host_emulation"):
        raise SystemExit("bench.py measures the CUDA library on a B200; the host-emulation build of the kernels is test infrastructure")
    if a.impl == "reference":
        if rank != 0:
            return
        # the sample is cut from the same generated matrix; generation needs the GPU only as a data source
        d = generate(sbmf, a.workload, 0)
        sample = take_sample(d, a.cpu_sample)
        del d
        pairs = max(1, (a.steps + 1) // 2)
        cb = cpu_reference(sample, K, pairs=pairs, port_steps=max(1, a.steps))
        cfg["n_train_sample"] = int(sample["train_user"].size)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": cb["sweeps_timed"],
                "warmup": 1, "ms_per_step": cb["s_per_sweep"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return

    dist = None
    nccl_id = None
    if world > 1:   # control plane: torch.distributed (gloo) for the id broadcast, barriers and the max over ranks
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo")

        def new_id():
            t = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                t = torch.frombuffer(bytearray(sbmf.nccl_unique_id()), dtype=torch.uint8).clone()
            dist.broadcast(t, 0)
            return t.numpy().tobytes()

        def max_over_ranks(x):
            t = torch.tensor([float(x)], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        nccl_id = new_id()
    else:
        def max_over_ranks(x):
            return float(x)

    def barrier():
        if dist is not None:
            dist.barrier()

    def mk(**kw):
        if world > 1:
            kw.update(rank=rank, world_size=world, nccl_id=kw.pop("nccl_id"))
        else:
            kw.pop("nccl_id", None)
        return sbmf.SbmfModel(**kw)

    cfg["parallelism"] = f"{world} GPU(s): users (CSR) and items (CSC) sharded by rating count, factors replicated" if world > 1 else "1 GPU"
    dev = local_rank
    t0 = time.perf_counter()
    d = generate(sbmf, a.workload, dev)
    gen_s = time.perf_counter() - t0
    n_train, n_test = int(d["train_user"].size), int(d["test_user"].size)
    cfg.update({"n_train": n_train, "n_test": n_test, "synth_seconds": round(gen_s, 2)})
    fu_per_sweep = float(n_train) * K

    # ---- value: K sweeps, everything resident, CUDA events on the library's stream
    m = mk(K=K, device=dev, sample_mode=sbmf.SAMPLE_REF, seed=1, nccl_id=nccl_id)
    m.set_train(d["train_user"], d["train_item"], d["train_rating"], I, J)
    m.set_test(d["test_user"], d["test_item"], d["test_rating"])
    m.init_factors()
    m.set_timing_enabled(0)
    m.sweep(W)
    m.synchronize()
    m.reset_timing()
    barrier()
    with ClockSampler(dev) as cs:
        tw0 = time.perf_counter()
        m.sweep(a.steps)
        ms = m.last_sweep_call_ms()     # waits for the K sweeps; device time between two events on the stream
        wall_ms = (time.perf_counter() - tw0) * 1e3
    barrier()
    ms = max_over_ranks(ms)             # max over ranks of the device time
    wall_ms = max_over_ranks(wall_ms)
    launches = m.timing()["kernel_launches"]
    clocks = cs.summary()
    ms_per_step = ms / a.steps
    value = fu_per_sweep / (ms_per_step * 1e-3)
    rmse = m.eval()[0]

    # ---- roofline of the dominant kernel: per-launch CUDA events (detail timing), 3 more sweeps
    m.set_timing_enabled(2)
    m.reset_timing()
    m.sweep(3)
    t = m.timing()
    peak, peak_src = peaks()
    roof = None
    t_local_ratings = t["top_kernel_ratings"]
    if t["top_kernel_launches"]:
        us = t["ms_top_kernel"] / t["top_kernel_launches"] * 1e3
        alg_bytes = 12.0 * 8 * t["top_kernel_ratings"]          # 12 B per (rating, dimension) visit x 8 dimensions per launch
        ach = alg_bytes / (us * 1e-6) / 1e9
        roof = {"kernel": "heavy_accumulate_kernel<2,2> (item-phase streaming block step)", "bound": "hbm", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "us_per_launch": us, "launches_timed": int(t["top_kernel_launches"]),
                "units_per_launch": f"{int(t['top_kernel_ratings'])} ratings x 8 dimensions", "algorithmic_bytes_per_launch": alg_bytes,
                "traffic": None,
                "note": "algorithmic bytes follow SURVEY 8(d) (12 B per rating x DIMENSION visit, the reference's per-dimension formulation); "
                        "the Gram-blocked kernel touches e/idx once per 8 dimensions, so frac > 1 is expected -- its real bound is the "
                        "L1TEX gather rate (one 32 B sector per clock per SM), see DESIGN.md"}
        prof = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.exists(prof):
            roof["traffic"] = json.load(open(prof)).get("heavy_accumulate_kernel<2,2>", {}).get("dram_bytes_per_launch")
            if roof["traffic"]:   # what the kernel really moves through HBM (ncu dram bytes of one launch) at the live launch time
                roof["dram_achieved_gbs"] = roof["traffic"] / (us * 1e-6) / 1e9
                roof["dram_frac"] = roof["dram_achieved_gbs"] / peak
        # the bound this kernel actually runs against: random 32-byte sector gathers from an L2-resident table, one LDG.E.256 per
        # lane.  tools/gather_probe.cu measures that access form alone on B200: 0.88-0.90 sectors per clock per SM (8.2-8.3 TB/s).
        try:
            props_sms, sm_mhz = 148, 1965.0
            sectors = 2.0 * t["top_kernel_ratings"]             # previous block (delta apply) + current block (accumulate) per rating
            spc = sectors / (us * 1e-6) / (props_sms * sm_mhz * 1e6)
            roof["gather_path"] = {"sectors_per_launch": sectors, "achieved_sectors_per_clk_per_sm": spc, "probe_peak_sectors_per_clk_per_sm": 0.90,
                                   "frac": spc / 0.90, "assumes": f"{props_sms} SMs at {sm_mhz:.0f} MHz (see clocks)",
                                   "source": "tools/gather_probe.cu (profiles/gather_probe_r1.txt)"}
        except Exception:   # noqa: BLE001 - informational only
            pass
    phases = {k: round(t[k] / max(t["sweeps"], 1), 3) for k in ("ms_rebuild", "ms_hypers", "ms_user_phase", "ms_exchange", "ms_item_phase", "ms_eval", "ms_allgather", "ms_total")}
    sweep_roof = {"algorithmic_bytes_per_sweep": ALG_BYTES_PER_FU * fu_per_sweep, "achieved_gbs_per_gpu": ALG_BYTES_PER_FU * value / 1e9 / world,
                  "frac_of_peak": ALG_BYTES_PER_FU * value / 1e9 / peak / world}
    m.close()

    # ---- e2e: the whole job through the C ABI from pinned host buffers
    e2e = None
    if not a.no_e2e:
        hb = {}
        for k in ("train_user", "train_item", "train_rating", "test_user", "test_item", "test_rating"):
            hb[k] = sbmf.pinned_empty(d[k].size, d[k].dtype)
            hb[k][:] = d[k]
        pred = sbmf.pinned_empty(n_test, np.float32)
        m2 = mk(K=K, device=dev, sample_mode=sbmf.SAMPLE_REF, seed=1, nccl_id=new_id() if world > 1 else None)
        m2.set_timing_enabled(0)
        m2.synchronize()
        barrier()
        te0 = time.perf_counter()
        m2.set_train(hb["train_user"], hb["train_item"], hb["train_rating"], I, J)
        te1 = time.perf_counter()
        m2.set_test(hb["test_user"], hb["test_item"], hb["test_rating"])
        m2.init_factors()
        te2 = time.perf_counter()
        last = None
        for _ in range(a.steps):
            m2.sweep(1)
            last = m2.eval()            # D2H read of the step's result (2 doubles), synchronises
        te3 = time.perf_counter()
        m2._ck(m2.lib.sbmf_cuda_get_pred(m2.h, pred.ctypes.data))
        m2.synchronize()
        barrier()
        te4 = time.perf_counter()
        e2e_s = max_over_ranks(te4 - te0)
        breakdown = {"set_train_s": round(te1 - te0, 4), "set_test_init_s": round(te2 - te1, 4), "sweeps_s": round(te3 - te2, 4),
                     "get_pred_s": round(te4 - te3, 4)}
        h2d = 12.0 * (n_train + n_test)
        d2h = 16.0 * a.steps + 4.0 * n_test
        e2e = {"value": fu_per_sweep * a.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / a.steps, "d2h_bytes_per_step": d2h / a.steps,
               "seconds_total": e2e_s, "sweeps": a.steps, "final_rmse": last[0], "breakdown_rank0": breakdown,
               "what": "set_train(H2D COO + device CSR/CSC build) + set_test + init_factors + steps x (sweep + eval D2H) + get_pred D2H, wall clock"}
        m2.close()

    if rank != 0:
        return
    cb = None
    if not a.no_cpu_baseline and world == 1:
        cb = cpu_reference(take_sample(d, a.cpu_sample), K, pairs=1)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "sweeps_per_s": 1e3 / ms_per_step, "wall_ms_per_step": wall_ms / a.steps, "rmse_after_timed": rmse, "clocks": clocks,
            "gpu_launches": int(launches), "phases_ms": phases, "roofline": roof, "roofline_sweep": sweep_roof, "e2e": e2e, "cpu_baseline": cb,
            "paper_i5_openmp_fu_per_s": 25.4e6}
    emit(line)


if __name__ == "__main__":
    main()
