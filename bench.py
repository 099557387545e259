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


def generate(sbmf, workload, device, rank=0, world=1, barrier=None):
    I, J, NTRAIN, _ = WORKLOADS[workload]
    n = int(round(NTRAIN / (1 - TEST_FRAC)))
    if float(I) * float(J) > 1e11:
        if world == 1:
            return sbmf.synth_generate_host(I, J, n, test_frac=TEST_FRAC, seed=SEED)
        # the host sampler uses every host thread: one rank generates, the others map the arrays from /dev/shm
        keys = ("train_user", "train_item", "train_rating", "test_user", "test_item", "test_rating")
        base = f"/dev/shm/sbmf_bench_{workload}_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            d = sbmf.synth_generate_host(I, J, n, test_frac=TEST_FRAC, seed=SEED)
            for k in keys:
                np.save(f"{base}_{k}.npy", d[k])
        barrier()
        if rank != 0:
            d = {k: np.load(f"{base}_{k}.npy", mmap_mode="r") for k in keys}
            d = {k: np.ascontiguousarray(v) for k, v in d.items()}
            d["num_users"], d["num_items"] = I, J
        barrier()
        if rank == 0:
            for k in keys:
                os.remove(f"{base}_{k}.npy")
        return d
    return sbmf.synth_generate(I, J, n, test_frac=TEST_FRAC, seed=SEED, device=device)
TEST_FRAC = 0.1
SEED = 20151001 + 3
METRIC = "gibbs_factor_updates_per_s"
UNIT = "factor-updates/s"
ALG_BYTES_PER_FU = 24.0          # 12 B per (rating, dimension) visit and half-step, two half-steps
PKG = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200")
FULL_POINT = "ml10m_k100"        # BASELINE.json configs[1]: small enough for the reference's CPU program to run the WHOLE config


def workload_config(name, world):
    """The `config` object of the JSON line: a pure function of (workload, N), so that our arm and the reference arm print the
    same object -- everything measured at run time (actual rating counts, generator seconds, ...) goes under "run"."""
    I, J, NTRAIN, K = WORKLOADS[name]
    return {"workload": f"{name}: synthetic Netflix/MovieLens-shaped Zipf rating matrix {I}x{J}, ~{NTRAIN} train ratings (+{int(TEST_FRAC * 100)}% test), K={K}; "
                        "one step = one full Gibbs sweep (gibbs_sbpmf2.cpp:335-637); the reference arm / cpu_baseline time the reference's own "
                        "program on a bounded sample of this matrix family (whole users, ~2M train ratings, all items, same K) -- see cpu_baseline.sample",
            "users": I, "items": J, "K": K, "l2_policy": "working set per sweep (>2 GB) far exceeds the 126 MB L2; no flush needed",
            "rebuild_every": 1, "residual_mode": "0: per-sweep residual rebuild of gibbs_sbpmf2.cpp:342-359 fused into the user phase",
            "sample_mode": "reference (x = mu + (1/lambda) z)",
            "parallelism": f"{world} GPU(s): users (CSR) and items (CSC) sharded by rating count, factors replicated" if world > 1 else "1 GPU"}


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
# `make -C oracle ref`), else the oracle port.  A bounded sample of the workload: the first users of the same matrix family.
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


def ref_tools():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from run_ref import ref_binary, run_ref
    return ref_binary, run_ref


def time_reference_files(train, test, K, threads, want_sweeps, timeout=900):
    """Per-sweep seconds of the UNMODIFIED reference program on triple files: (wall(T=Tb) - wall(T=1)) / (Tb - 1), so reading and
    indexing the files (identical in both runs) cancels.  Tb = the largest prebuilt sweep count with Tb - 1 <= want_sweeps."""
    ref_binary, run_ref = ref_tools()
    b1 = ref_binary(K, 1)
    Tb = next((t for t in (21, 11, 6, 3, 2) if t - 1 <= max(want_sweeps, 1) and ref_binary(K, t)), None)
    if not b1 or not Tb:
        return None
    r1 = run_ref(b1, train, test, threads=threads, timeout=timeout)
    rb = run_ref(ref_binary(K, Tb), train, test, threads=threads, timeout=timeout)
    if len(r1["rmse"]) != 1 or len(rb["rmse"]) != Tb:
        return None
    return {"s_per_sweep": (rb["wall_s"] - r1["wall_s"]) / (Tb - 1), "sweeps_timed": Tb - 1, "rmse": rb["rmse"], "load_s": r1["wall_s"] - (rb["wall_s"] - r1["wall_s"]) / (Tb - 1),
            "num_rows": rb["num_rows"], "num_users": rb["num_users"], "num_items": rb["num_items"]}


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


def cpu_reference_on_files(train, test, n_train, K, desc, want_sweeps, all_threads=True):
    """cpu_baseline object from triple files: the reference binary at 1 thread and (optionally) at all host threads; best one wins."""
    ncpu = os.cpu_count() or 1
    r1 = time_reference_files(train, test, K, 1, want_sweeps)
    if not r1:
        return None
    runs = {1: r1}
    if all_threads and ncpu > 1:
        try:
            rn = time_reference_files(train, test, K, ncpu, min(want_sweeps, 2), timeout=max(120.0, 40 * r1["s_per_sweep"] + 60))
            if rn:
                runs[ncpu] = rn
        except Exception:   # noqa: BLE001 - glibc rand() lock contention can make the OpenMP run pathologically slow
            desc += f"; OMP_NUM_THREADS={ncpu} run timed out (rand() lock contention, SURVEY.md 0.8)"
    best = min(runs, key=lambda t: runs[t]["s_per_sweep"])
    return {"value": n_train * K / runs[best]["s_per_sweep"], "unit": UNIT, "cores": best, "kind": "reference",
            "sample": desc + "; unmodified gibbs_sbpmf2.cpp (g++ -O3 -fopenmp), per-sweep time = (wall(T=1+n) - wall(T=1)) / n so file loading cancels",
            "by_threads": {str(t): n_train * K / v["s_per_sweep"] for t, v in runs.items()}, "host_cpus": ncpu,
            "sweeps_timed": runs[best]["sweeps_timed"], "s_per_sweep": runs[best]["s_per_sweep"], "load_s": runs[best]["load_s"],
            "n_train": int(n_train)}


def cpu_reference(sample, K, want_sweeps=2):
    """cpu_baseline of OUR arm: the reference binary on a sample cut from the matrix this run generated (port if the binaries are absent)."""
    n = int(sample["train_user"].size)
    desc = f"first {sample['num_users']} users of the workload ({n} train ratings, {sample['num_items']} items, K={K})"
    out = None
    try:
        with tempfile.TemporaryDirectory(prefix="sbmf_bench_") as tmp:
            tr, te = os.path.join(tmp, "train"), os.path.join(tmp, "test")
            write_triples(tr, sample["train_user"], sample["train_item"], sample["train_rating"])
            # the reference sizes its arrays by max id over train U test: make the item id space explicit with one test row
            write_triples(te, np.append(sample["test_user"], sample["num_users"] - 1), np.append(sample["test_item"], sample["num_items"] - 1),
                          np.append(sample["test_rating"], 3.0))
            out = cpu_reference_on_files(tr, te, n, K, desc, want_sweeps)
    except Exception as e:   # noqa: BLE001 - the reference binary may be absent or time out; fall back to the port
        desc += f"; reference binary failed: {type(e).__name__}"
    if out is None:
        r = run_oracle_port(sample, K, 2, 1)
        out = {"value": n * K / r["s_per_sweep"], "unit": UNIT, "cores": 1, "kind": "port",
               "sample": desc + "; oracle/sbmf_oracle.c (scalar C restatement of gibbs_sbpmf2.cpp), oracle/_ref binaries absent",
               "host_cpus": os.cpu_count() or 1, "sweeps_timed": r["sweeps_timed"], "s_per_sweep": r["s_per_sweep"]}
    return out


def synth_files(workload, tmp, max_train):
    """bin/sbmf_synth (host generator, no CUDA, no libsbmf_cuda.so) -> triple files of (a whole-user prefix of) the workload."""
    I, J, NTRAIN, _ = WORKLOADS[workload]
    n = int(round(NTRAIN / (1 - TEST_FRAC)))
    tr, te = os.path.join(tmp, f"{workload}.train"), os.path.join(tmp, f"{workload}.test")
    exe = os.path.join(PKG, "bin", "sbmf_synth")
    r = subprocess.run([exe, "-users", str(I), "-items", str(J), "-ratings", str(n), "-seed", str(SEED), "-test_frac", str(TEST_FRAC),
                        "-max_train", str(max_train), "-train", tr, "-test", te], capture_output=True, text=True, check=True)
    return tr, te, json.loads(r.stdout.strip().splitlines()[-1])


def reference_arm(a, world):
    """bench.py --impl reference: the reference's own CPU program on the box's host cores.  This process loads NO library of this
    repository: the sample comes from the stand-alone host generator (bin/sbmf_synth), the timed thing is a subprocess of the
    unmodified reference binary.  Falls back to the oracle port (a C restatement) only if oracle/_ref was not built."""
    I, J, NTRAIN, K = WORKLOADS[a.workload]
    cfg = workload_config(a.workload, world)
    with tempfile.TemporaryDirectory(prefix="sbmf_ref_arm_") as tmp:
        t0 = time.perf_counter()
        tr, te, info = synth_files(a.workload, tmp, a.cpu_sample)
        gen_s = time.perf_counter() - t0
        full = info["n_train"] == info["n_train_full"]
        desc = (f"{'the WHOLE workload' if full else 'first ' + str(info['num_users']) + ' users of the workload'} ({info['n_train']} of {info['n_train_full']} train ratings, "
                f"{info['num_items']} items, K={K}), matrix from the stand-alone host generator bin/sbmf_synth (same family, shape, seed and Zipf exponents as the GPU arm's device generator)")
        cb = cpu_reference_on_files(tr, te, info["n_train"], K, desc, a.steps)
        if cb is None:   # no reference binaries on this box: the oracle port
            u, i, r = np.loadtxt(tr, dtype=np.float64, ndmin=2).T
            su, si, sr = np.loadtxt(te, dtype=np.float64, ndmin=2).T
            sample = {"train_user": u.astype(np.uint32), "train_item": i.astype(np.uint32), "train_rating": r.astype(np.float32), "test_user": su.astype(np.uint32),
                      "test_item": si.astype(np.uint32), "test_rating": sr.astype(np.float32), "num_users": info["num_users"], "num_items": info["num_items"]}
            rr = run_oracle_port(sample, K, max(1, min(a.steps, 5)), 1)
            cb = {"value": info["n_train"] * K / rr["s_per_sweep"], "unit": UNIT, "cores": 1, "kind": "port", "sample": desc + "; oracle/sbmf_oracle.c, oracle/_ref binaries absent",
                  "host_cpus": os.cpu_count() or 1, "sweeps_timed": rr["sweeps_timed"], "s_per_sweep": rr["s_per_sweep"], "n_train": info["n_train"]}
        # one point where the reference runs the COMPLETE configuration (BASELINE.json configs[1], ML-10M-shaped K=100): the GPU
        # arm prints the same key for the same workload, so this pair is a same-config ratio
        full_point = None
        if not a.no_full_point and a.workload != FULL_POINT:
            try:
                ftr, fte, finfo = synth_files(FULL_POINT, tmp, 0)
                fK = WORKLOADS[FULL_POINT][3]
                fr = time_reference_files(ftr, fte, fK, 1, 2, timeout=1200)
                if fr:
                    full_point = {"workload": FULL_POINT, "config": workload_config(FULL_POINT, 1), "value": finfo["n_train"] * fK / fr["s_per_sweep"], "unit": UNIT,
                                  "s_per_sweep": fr["s_per_sweep"], "sweeps_timed": fr["sweeps_timed"], "n_train": finfo["n_train"], "cores": 1,
                                  "load_s": fr["load_s"], "what": "unmodified gibbs_sbpmf2.cpp on the complete ML-10M-shaped K=100 configuration, 1 thread"}
            except Exception as e:   # noqa: BLE001 - informational
                full_point = {"workload": FULL_POINT, "error": f"{type(e).__name__}: {e}"}
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": cb["sweeps_timed"], "steps_requested": a.steps,
            "warmup": a.warmup, "ms_per_step": cb["s_per_sweep"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "run": {"n_train_sample": info["n_train"], "n_train_workload": info["n_train_full"], "synth_seconds": round(gen_s, 2),
                    "warmup_note": "the T=1 run of the pair is the warm-up: its sweep and the file loading are subtracted"},
            "full_config_point": full_point}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def load_traffic(want_ncu=False):
    """measured DRAM bytes per rating of the dominant kernel (ncu --set full, profiles/): newest round first"""
    for name in ("traffic_r2.json", "traffic_r1.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            t = json.load(open(p)).get("heavy_accumulate_kernel<2,2>", {})
            if want_ncu:
                return {k: t[k] for k in ("l1tex_data_pipe_lsu_wavefronts_pct", "lts_t_sectors_srcunit_tex_pct", "l1tex_sectors_global_ld", "duration_us_under_ncu", "source") if k in t}
            if t.get("dram_bytes_per_rating"):
                return float(t["dram_bytes_per_rating"]), name
            if t.get("dram_bytes_per_launch") and t.get("ratings_per_launch"):
                return float(t["dram_bytes_per_launch"]) / float(t["ratings_per_launch"]), name
    return {} if want_ncu else (None, None)


def build_roofline(t, pr, peak_hbm, peak_src, K, KB, ms_per_step, n_test, world):
    """Per launch of the dominant kernel and per sweep: streamed bytes against the HBM peak AND 32-byte sectors delivered by L2
    against the gather rates sbmf_cuda_probe measured on this device a moment ago.  `bound` names the larger fraction."""
    forms = pr["gather_sectors_per_s"]
    best_form = max(forms, key=lambda k: forms[k])
    same, best = forms["g32_lane"], forms[best_form]
    roof = None
    R = t["top_kernel_ratings"]
    if t["top_kernel_launches"] and R:
        us = t["ms_top_kernel"] / t["top_kernel_launches"] * 1e3
        sec = us * 1e-6
        sectors = 2.0 * R                       # block b-1 (delta apply) + block b (accumulate): one 32-byte sector each per rating
        l2 = {"sectors_per_launch": sectors, "achieved_gbs": sectors * 32 / sec / 1e9, "achieved_sectors_per_clk_per_sm": sectors / sec / (pr["sm_count"] * pr["sm_clock_mhz_max"] * 1e6),
              "peak_same_form_gbs": same * 32 / 1e9, "peak_best_form_gbs": best * 32 / 1e9, "best_form": best_form,
              "frac_same_form": sectors / sec / same, "frac_best_form": sectors / sec / best,
              "what": "factor-row gathers from the L2-resident K8 block (2 per rating); peaks = sbmf_cuda_probe on this device in this run: "
                      "same_form = one random 32-byte sector per lane (how the kernel gathers), best_form = the fastest access form probed"}
        stream = 12.0 * R                       # idx 4 B + e read 4 B + e write 4 B per rating per launch
        hbm = {"algorithmic_bytes_per_launch": stream, "achieved_gbs": stream / sec / 1e9, "peak_gbs": peak_hbm, "peak_source": peak_src,
               "peak_probe_copy_gbs": pr["hbm_copy_gbs"], "frac": stream / sec / 1e9 / peak_hbm}
        per_rating, tsrc = load_traffic()
        if per_rating:
            hbm["traffic_bytes_per_rating"] = per_rating
            hbm["traffic_source"] = f"profiles/{tsrc} (ncu dram__bytes_read+write of one launch / its ratings), scaled to this rank's {int(R)} ratings"
        bound_l2 = l2["frac_best_form"] >= hbm["frac"]
        survey = 12.0 * 8 * R / sec / 1e9 / peak_hbm
        roof = {"kernel": "heavy_accumulate_kernel<2,2> (item-phase streaming block step)", "bound": "l2_gather" if bound_l2 else "hbm",
                "achieved": l2["achieved_gbs"] if bound_l2 else hbm["achieved_gbs"], "peak": l2["peak_best_form_gbs"] if bound_l2 else peak_hbm,
                "unit": "GB/s", "frac": l2["frac_best_form"] if bound_l2 else hbm["frac"],
                "peak_source": "sbmf_cuda_probe (csrc/probe.cu), same device, same run" if bound_l2 else peak_src,
                "traffic": per_rating * R if per_rating else None, "us_per_launch": us, "launches_timed": int(t["top_kernel_launches"]),
                "units_per_launch": f"{int(R)} ratings x 8 dimensions (this rank)", "hbm": hbm, "l2_gather": l2,
                "frac_survey_8d": survey,
                "ncu_committed": load_traffic(want_ncu=True),
                "note": "the binding unit is the SM's L1TEX data pipe: a scattered 32-byte sector costs one wavefront there (ncu_committed: "
                        "pipe utilisation of this kernel under ncu; the live figure is l2_gather: sectors delivered per second against the "
                        "rates sbmf_cuda_probe measured in this run -- frac_same_form above 1 means adjacent lanes share lines on the popular "
                        "rows, which the probe's random ids never do).  frac_survey_8d prices the launch by SURVEY 8(d) (12 B per rating x "
                        "DIMENSION, the reference's per-dimension formulation); the Gram-blocked kernel streams idx/e once per 8 dimensions, so "
                        "that figure exceeds 1 and is kept only for comparison"}
    # whole sweep: sector gathers issued by the phases + evaluation against the same probed rates
    sect_sweep = KB * (t["nnz_light_user"] + 2.0 * t["nnz_heavy_user"] + t["nnz_light_item"] + 2.0 * t["nnz_heavy_item"]) + 2.0 * KB * n_test / world
    s = ms_per_step * 1e-3
    sweep = {"l2_sectors_per_sweep": sect_sweep, "l2_achieved_gbs": sect_sweep * 32 / s / 1e9, "l2_frac_same_form": sect_sweep / s / same,
             "l2_frac_best_form": sect_sweep / s / best,
             "what": "K/8 sector gathers per rating and phase for register-resident rows, 2 K/8 for streamed rows (apply + accumulate), 2 K/8 per test pair; this rank"}
    return roof, sweep


def mgpu_parity(sbmf, dev, rank, world, new_id, max_over_ranks):
    """N > 1: the G-GPU chain against a 1-GPU chain of the same seed on a small workload (ML-1M-shaped, K = 50, live sampling,
    5 sweeps; draws are keyed by global row id, so the chains agree to fp32 summation order)."""
    s = sbmf.synth_generate(6040, 3706, 1000209, seed=20151001, device=dev)
    outs = []
    for g in (True, False):
        kw = dict(K=50, device=dev, sample_mode=sbmf.SAMPLE_REF, seed=11)
        if g:
            kw.update(rank=rank, world_size=world, nccl_id=new_id())
        m = sbmf.SbmfModel(**kw)
        m.set_train(s["train_user"], s["train_item"], s["train_rating"], 6040, 3706)
        m.set_test(s["test_user"], s["test_item"], s["test_rating"])
        m.init_factors()
        m.set_timing_enabled(0)
        m.sweep(5)
        outs.append((m.get_state(with_E=False), m.rmse_history(0, 5)[0].copy()))
        m.close()

    def rel(a, b):
        return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / max(float(np.max(np.abs(b))), 1e-30))
    (a, ra), (b, rb) = outs
    res = {k: max_over_ranks(rel(a[k], b[k])) for k in ("U", "V", "b_i", "b_j")}
    res["rmse_history"] = max_over_ranks(float(np.max(np.abs(ra - rb))))
    return {"workload": "ml1m-shaped synthetic 6040x3706, 1.0M ratings, K=50, live sampling, 5 sweeps", "against": "1-GPU chain of the same seed on each rank's own GPU",
            "max_rel_diff": res, "ok": bool(max(res.values()) <= 1e-4)}


def full_config_point(sbmf, dev):
    """The GPU side of the reference arm's full_config_point: the complete ML-10M-shaped K=100 configuration on one GPU."""
    I, J, _, K = WORKLOADS[FULL_POINT]
    d = generate(sbmf, FULL_POINT, dev)
    m = sbmf.SbmfModel(K=K, device=dev, sample_mode=sbmf.SAMPLE_REF, seed=1)
    m.set_train(d["train_user"], d["train_item"], d["train_rating"], I, J)
    m.set_test(d["test_user"], d["test_item"], d["test_rating"])
    m.init_factors()
    m.set_timing_enabled(0)
    m.sweep(5)
    m.synchronize()
    m.sweep(20)
    ms = m.last_sweep_call_ms() / 20
    n = int(d["train_user"].size)
    m.close()
    return {"workload": FULL_POINT, "config": workload_config(FULL_POINT, 1), "value": n * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "sweeps_timed": 20,
            "n_train": n, "what": "the complete ML-10M-shaped K=100 configuration on one B200, everything resident, CUDA events"}


def e2e_cli(workload, steps, ref_load_s=None, ref_load_n=None):
    """The drop-in job FROM FILES: bin/sbmf -train F -test F -dim 1,1,K -iter steps -out P on the workload written as triple text
    and as libFM binary (.x/.y) by bin/sbmf_synth; stage seconds from the CLI's own -timing file (parse = the reader of
    csrc/rating_reader.h replacing the three getline + sscanf passes of gibbs_sbpmf2.cpp:32-221; build = set_train + set_test)."""
    I, J, NTRAIN, K = WORKLOADS[workload]
    n = int(round(NTRAIN / (1 - TEST_FRAC)))
    exe, synth = os.path.join(PKG, "bin", "sbmf"), os.path.join(PKG, "bin", "sbmf_synth")
    out = {"workload": workload, "what": "bin/sbmf from files: parse + device build + sweeps (each followed by the RMSE read-back and its `rmse is` line) + -out predictions; wall clock of the whole process and the CLI's own stage timers"}
    with tempfile.TemporaryDirectory(prefix="sbmf_cli_") as tmp:
        for fmt in ("triples", "libfm_binary"):
            tr, te = os.path.join(tmp, f"{fmt}.train"), os.path.join(tmp, f"{fmt}.test")
            t0 = time.perf_counter()
            g = subprocess.run([synth, "-users", str(I), "-items", str(J), "-ratings", str(n), "-seed", str(SEED), "-test_frac", str(TEST_FRAC),
                                "-binary", "1" if fmt == "libfm_binary" else "0", "-train", tr, "-test", te], capture_output=True, text=True, check=True)
            info = json.loads(g.stdout.strip().splitlines()[-1])
            write_s = time.perf_counter() - t0
            tj = os.path.join(tmp, f"{fmt}.timing.json")
            t0 = time.perf_counter()
            r = subprocess.run([exe, "-train", tr, "-test", te, "-dim", f"1,1,{K}", "-iter", str(steps), "-seed", "1", "-out", os.path.join(tmp, f"{fmt}.pred"),
                                "-timing", tj], capture_output=True, text=True, cwd=tmp)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                out[fmt] = {"error": (r.stderr or r.stdout)[-300:]}
                continue
            st = json.load(open(tj))
            rm = [float(l.split()[-1]) for l in r.stdout.splitlines() if l.startswith("rmse is")]
            files = [tr, te] if fmt == "triples" else [tr + ".x", tr + ".y", te + ".x", te + ".y"]
            out[fmt] = {"value": info["n_train"] * K * steps / wall, "unit": UNIT, "wall_s": wall, "stages": st, "final_rmse": rm[-1] if rm else None,
                        "input_bytes": int(sum(os.path.getsize(f) for f in files)), "n_train": info["n_train"], "files_written_in_s": round(write_s, 2)}
            for f in files:
                os.remove(f)
    if ref_load_s is not None:
        out["reference_load"] = {"seconds": ref_load_s, "n_train": ref_load_n, "ratings_per_s": ref_load_n / ref_load_s if ref_load_s > 0 else None,
                                 "what": "the reference's own three getline + sscanf passes + jagged-array build (gibbs_sbpmf2.cpp:32-221) on the cpu_baseline sample: wall(T=1) minus one sweep"}
    return out


# ---------------------------------------------------------------------------------------------------------------------
# general FM Gibbs (csrc/fm.cu behind include/sbmf_fm_cuda.h): libFM's -method mcmc on a design matrix
FM_WORKLOADS = {
    # name: (users, items, train ratings, K, dense real-valued context attributes per case)
    "fm_mf_ml10m": (71567, 10681, 10000000, 8, 0),      # matrix factorisation as an FM: one-hot user + one-hot item
    "fm_wide_ml1m": (6040, 3706, 1000209, 8, 4),        # ... plus 4 dense attributes (every column of them has one entry per case)
}


def fm_bench(a):
    """python bench.py --workload fm_mf_ml10m | fm_wide_ml1m [--impl reference]: iterations/s of libFM's MCMC learner
    (fm_learn_mcmc.h:411-623, 780-835).  Ours: sbmf_fm_learn on one B200.  Reference: the unmodified libFM built from
    /root/reference (oracle/_ref/libFM -method mcmc) on the same files, one thread; per-iteration time = (wall(iter=3) - wall(iter=1)) / 2."""
    I, J, NTRAIN, K, W = FM_WORKLOADS[a.workload]
    n = int(round(NTRAIN / (1 - TEST_FRAC)))
    cfg = {"workload": f"{a.workload}: synthetic MovieLens-shaped ratings {I}x{J}, ~{NTRAIN} train cases, cast as a factorization machine "
                       f"(one-hot user + one-hot item{' + %d dense real-valued attributes' % W if W else ''}), K={K}, libFM -method mcmc -task r",
           "users": I, "items": J, "K": K, "dense_attributes": W, "parallelism": "1 GPU"}
    metric, unit = "fm_gibbs_iterations_per_s", "iterations/s"
    with tempfile.TemporaryDirectory(prefix="sbmf_fm_") as tmp:
        tr, te = os.path.join(tmp, "fm.train"), os.path.join(tmp, "fm.test")
        subprocess.run([os.path.join(PKG, "bin", "sbmf_synth"), "-users", str(I), "-items", str(J), "-ratings", str(n), "-seed", str(SEED), "-test_frac", str(TEST_FRAC),
                        "-libfm_text", "1", "-train", tr, "-test", te], capture_output=True, text=True, check=True)

        def dense(nrows, seed):
            rs = np.random.RandomState(seed)
            return (np.round(rs.standard_normal((nrows, W)) * 8) / 8).astype(np.float32)

        if W:   # append the dense attributes to every line (ids after users and items); 1M-line files: plain Python is fine
            for path, seed in ((tr, 1), (te, 2)):
                lines = open(path).read().splitlines()
                x = dense(len(lines), seed)
                with open(path, "w") as f:
                    for ln, row in zip(lines, x):
                        f.write(ln + "".join(f" {I + J + k}:{row[k]:g}" for k in range(W)) + "\n")
        if a.impl == "reference":
            exe = os.path.join(ROOT, "oracle", "_ref", "libFM")
            if not os.path.exists(exe):
                emit({"impl": "reference", "unavailable": "oracle/_ref/libFM not built (needs /root/reference at build time)"})
                return
            walls = {}
            for it in (1, 3):
                t0 = time.perf_counter()
                r = subprocess.run([exe, "-task", "r", "-train", tr, "-test", te, "-dim", f"1,1,{K}", "-iter", str(it), "-method", "mcmc", "-init_stdev", "0.1"],
                                   capture_output=True, text=True, cwd=tmp, env=dict(os.environ, OMP_NUM_THREADS="1"))
                walls[it] = time.perf_counter() - t0
                if r.returncode != 0:
                    emit({"impl": "reference", "unavailable": "libFM failed: " + (r.stderr or r.stdout)[-200:]})
                    return
            per = (walls[3] - walls[1]) / 2
            rm = [float(x.split("Test=")[1]) for x in r.stdout.splitlines() if x.startswith("#Iter")]
            cb = {"value": 1 / per, "unit": unit, "cores": 1, "kind": "reference", "sample": "the whole workload; unmodified libFM 1.4.2 of the reference (src/libfm/libfm.cpp, g++ -O3), -method mcmc",
                  "s_per_iteration": per, "load_s": walls[1] - per, "test_rmse": rm}
            emit({"impl": "reference", "metric": metric, "value": 1 / per, "unit": unit, "n_gpus": 1, "steps": 2, "warmup": 1, "ms_per_step": per * 1e3, "higher_is_better": True,
                  "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
                  "e2e": {"value": 1 / per, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
            return
        import sbmf
        d = sbmf.synth_generate_host(I, J, n, test_frac=TEST_FRAC, seed=SEED)   # the matrix bin/sbmf_synth wrote (same generator, same seed)

        def fm(u, i, r, seed):
            nr = u.size
            cols = [u.astype(np.uint32), (I + i).astype(np.uint32)] + [np.full(nr, I + J + k, dtype=np.uint32) for k in range(W)]
            x = dense(nr, seed) if W else None
            vals = [np.ones(nr, np.float32), np.ones(nr, np.float32)] + [x[:, k] for k in range(W)]
            return {"row_ptr": ((2 + W) * np.arange(nr + 1)).astype(np.int64), "attr": np.stack(cols, axis=1).reshape(-1), "x": np.stack(vals, axis=1).reshape(-1),
                    "y": r.astype(np.float32)}
        mtr, mte = fm(d["train_user"], d["train_item"], d["train_rating"], 1), fm(d["test_user"], d["test_item"], d["test_rating"], 2)
        p = I + J + W + 1     # libFM counts one attribute beyond the largest id it reads (libfm.cpp:326)
        group = np.zeros(p, np.uint32)
        m = sbmf.FmModel(p, K, attr_group=group, seed=1)
        t0 = time.perf_counter()
        m.set_train(mtr); m.set_test(mte); m.init()
        setup_s = time.perf_counter() - t0
        Wm = max(a.warmup, 3)
        m.learn(Wm); m.rmse_history(0, Wm)
        with ClockSampler(0) as cs:
            t0 = time.perf_counter()
            m.learn(a.steps)
            r = m.rmse_history(Wm, a.steps)          # synchronises
            dt = (time.perf_counter() - t0) / a.steps
        nnz = int(mtr["row_ptr"][-1])
        alg = nnz * (16.0 + K * 24.0 + K * 4.0 + 8.0 + 4.0 * K)   # w draw, v draws, q rebuild, re-prediction per entry and iteration
        peak, peak_src = peaks()
        roof = {"kernel": "fm_col_* / fm_predict_* (column passes of one iteration)", "bound": "hbm", "achieved": alg / dt / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg / dt / 1e9 / peak, "peak_source": peak_src, "traffic": None, "algorithmic_bytes_per_iteration": alg,
                "note": "whole-iteration figure (wall clock around sbmf_fm_learn + history read-back): 16 B per entry for the w draw, 24 B per entry and factor for "
                        "the v draws, 4 B per entry and factor for the q rebuild, 8 + 4K B per entry for libFM's full re-prediction"}
        emit({"metric": metric, "value": 1 / dt, "unit": unit, "n_gpus": 1, "steps": a.steps, "warmup": Wm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "f32 (column sums f64)", "data": "synthetic", "config": cfg, "clocks": cs.summary(), "roofline": roof,
              "run": {"cases": int(mtr["y"].size), "nnz": nnz, "attributes": p, "runs": int(m.get_runs().size - 1), "setup_s": round(setup_s, 3)},
              "test_rmse": [round(float(x), 5) for x in r[1]], "gpu_launches": None,
              "e2e": {"value": a.steps / (a.steps * dt + setup_s), "unit": unit, "h2d_bytes_per_step": (12.0 * nnz + 4.0 * mtr["y"].size) / a.steps, "d2h_bytes_per_step": 16.0,
                      "what": "set_train + set_test + init (H2D of the design matrix, device transpose) + the timed iterations"}})
        m.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netflix_k100", choices=sorted(WORKLOADS) + sorted(FM_WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=2000000, help="train ratings in the CPU baseline's sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-full-point", action="store_true", help="skip the ML-10M-shaped full-configuration point")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the G-GPU vs 1-GPU chain check")
    ap.add_argument("--no-cli", action="store_true", help="skip e2e_cli (the job from files through bin/sbmf)")
    ap.add_argument("--options", default="", help="name=value,... passed to sbmf_cuda_set_option on every handle (A/B measurements)")
    a = ap.parse_args()
    rank, local_rank, world = dist_env()
    if world != a.gpus and world != 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    if a.workload in FM_WORKLOADS:     # the general FM Gibbs path: one GPU (cases shard naturally; replicas only)
        if rank == 0:
            fm_bench(a)
        return
    I, J, NTRAIN, K = WORKLOADS[a.workload]
    W = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        if rank == 0:
            reference_arm(a, max(world, a.gpus))
        return
    cfg = workload_config(a.workload, world)
    opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in a.options.split(",") if kv}

    import sbmf
    if os.environ.get("SBMF_EMULATED") or hasattr(sbmf.load_library(), "sbmf_simt_host_emulation"):
        raise SystemExit("bench.py measures the CUDA library on a B200; the host-emulation build of the kernels is test infrastructure")

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
        new_id = None

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
        return sbmf.SbmfModel(options=opts, **kw)

    dev = local_rank
    t0 = time.perf_counter()
    d = generate(sbmf, a.workload, dev, rank, world, barrier)
    gen_s = time.perf_counter() - t0
    n_train, n_test = int(d["train_user"].size), int(d["test_user"].size)
    run = {"n_train": n_train, "n_test": n_test, "synth_seconds": round(gen_s, 2), "options": opts}
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

    # ---- roofline of the dominant kernel: per-launch CUDA events (detail timing), 3 more sweeps; then the probes on the same device
    m.set_timing_enabled(2)
    m.reset_timing()
    m.sweep(3)
    t = m.timing()
    m.synchronize()
    phases = {k: round(t[k] / max(t["sweeps"], 1), 3) for k in ("ms_rebuild", "ms_hypers", "ms_user_phase", "ms_exchange", "ms_item_phase", "ms_eval", "ms_allgather", "ms_total")}
    m.close()
    peak, peak_src = peaks()
    barrier()
    pr = sbmf.probe(dev)                # every rank probes its own GPU at the same time (none of them is running a sweep)
    barrier()
    roof, sweep_gather = build_roofline(t, pr, peak, peak_src, K, (K + 7) // 8, ms_per_step, n_test, world)
    sweep_roof = {"frac_survey_8d": ALG_BYTES_PER_FU * value / 1e9 / peak / world, "algorithmic_bytes_per_sweep_survey_8d": ALG_BYTES_PER_FU * fu_per_sweep,
                  "survey_8d_gbs_per_gpu": ALG_BYTES_PER_FU * value / 1e9 / world}
    sweep_roof.update(sweep_gather)
    probes = {"hbm_copy_gbs": pr["hbm_copy_gbs"], "hbm_read_gbs": pr["hbm_read_gbs"], "table_mb": pr["table_bytes"] / 2 ** 20,
              "gather_gbs": {k: v * 32 / 1e9 for k, v in pr["gather_sectors_per_s"].items()},
              "gather_sectors_per_clk_per_sm": {k: v / (pr["sm_count"] * pr["sm_clock_mhz_max"] * 1e6) for k, v in pr["gather_sectors_per_s"].items()},
              "sm_count": pr["sm_count"], "sm_clock_mhz_max": pr["sm_clock_mhz_max"], "source": "sbmf_cuda_probe (csrc/probe.cu): best of 5 launches per form, this device, this run"}

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
        set_train_max = max_over_ranks(te1 - te0)
        h2d = 12.0 * (n_train + n_test)
        d2h = 16.0 * a.steps + 4.0 * n_test
        e2e = {"value": fu_per_sweep * a.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d / a.steps, "d2h_bytes_per_step": d2h / a.steps,
               "seconds_total": e2e_s, "sweeps": a.steps, "final_rmse": last[0], "breakdown_rank0": breakdown, "set_train_s_max_over_ranks": round(set_train_max, 4),
               "what": "set_train(H2D COO + device CSR/CSC build) + set_test + init_factors + steps x (sweep + eval D2H) + get_pred D2H, wall clock"}
        m2.close()

    parity = None
    if world > 1 and not a.no_parity:
        parity = mgpu_parity(sbmf, dev, rank, world, new_id, max_over_ranks)

    if rank != 0:
        return
    cb = None
    fp = None
    if not a.no_cpu_baseline and world == 1:
        cb = cpu_reference(take_sample(d, a.cpu_sample), K)
    if not a.no_full_point and world == 1 and a.workload != FULL_POINT:
        fp = full_config_point(sbmf, dev)
    cli = None
    if not a.no_cli and world == 1:
        del d
        try:
            cli = e2e_cli(a.workload, a.steps, cb.get("load_s") if cb else None, cb.get("n_train") if cb else None)
        except Exception as e:   # noqa: BLE001 - an additional measurement: report, do not lose the line
            cli = {"error": f"{type(e).__name__}: {e}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "run": run,
            "sweeps_per_s": 1e3 / ms_per_step, "wall_ms_per_step": wall_ms / a.steps, "rmse_after_timed": rmse, "clocks": clocks,
            "gpu_launches": int(launches), "phases_ms": phases, "roofline": roof, "roofline_sweep": sweep_roof, "probes": probes, "e2e": e2e,
            "cpu_baseline": cb, "parity": parity, "full_config_point": fp, "e2e_cli": cli, "paper_i5_openmp_fu_per_s": 25.4e6}
    emit(line)


if __name__ == "__main__":
    main()
