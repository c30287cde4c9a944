#!/usr/bin/env python
"""bench.py — HOBBIT hot-path benchmark (one JSON line on stdout).

Workload (BASELINE.json configs[1], the largest single-GPU configuration the metric is quoted on):
    Our_PC commit_standard of a random multilinear polynomial with N = 2^26 coefficients, K = 32 chunks,
    linear-time mode: rows Reed–Solomon (NTT 2048 -> 4096), columns Orion expander code (n = 1024 -> 1761),
    column-wise BLAKE3 Merkle–Damgård leaves (2^21) + Merkle tree — i.e. `test_PC(1<<26, 4, 32)`'s commit.
A "step" is one whole commit of that polynomial.

    value : field-elems/s with the polynomial already resident in HBM (device pointer into the C ABI)
    e2e   : same call with the polynomial in pinned HOST memory (H2D inside the timed region, overlapped chunk by
            chunk) and every Merkle level copied back to the host (D2H) — what a reference-side caller sees
    roofline     : the dominant kernel (expander encode fused with the leaf hashing), CUDA-event timed per launch
    cpu_baseline : the unmodified reference (oracle/_ref, built from /root/reference) on a bounded sample

`--impl reference` times the reference's own CPU implementation of the same call on the host cores.
N > 1 (torchrun): ONE commitment of N*2^26 coefficients (32*N chunks of the same size) sharded over the ranks behind the C ABI
(hb_dist_commit_standard: chunks by rank, the encode kernel stores every inner leaf digest straight into the window of the rank that
owns the leaf range over NVLink, subtree levels scattered to every rank) -> weak scaling (fixed work per GPU); timing = max over ranks.
Outside the timed region the N > 1 run compares the sharded commitment with the single-GPU commitment of the same polynomial, level by
level, and a sharded sumcheck proof with the single-GPU proof ("parity_check").

"extras" (same JSON line, rank 0): the other BASELINE configurations, each run by the reference-free C++ tools through the host mirror
(hobbit_b200/pc_prove, hobbit_b200/mlp_prove): config 1 test_PC(2^20, 4, 32) commit + open seconds, configs 3/4 MLP / AES / SQL prove
seconds with ps_kb, and at N = 8 config 5 test_Elastic_PC(2^30, 2) commit + open with the stream in pinned host memory.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pc_commit_field_elems_per_s"
UNIT = "field-elems/s"
LOGN, K, TRS = 26, 32, 1024


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.rows = gpu, False, []

    def run(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.p.stdout:
                if self.stop_flag:
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        try:
            self.p.terminate()
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_poly(n, seed):
    """generate_randomness-shaped input (utils.cpp:873-883): real-only values c + rand(), c redrawn every 100 elements.
    Synthetic: drawn with numpy (the libc-RNG-exact generator is exercised in the parity tests)."""
    rng = np.random.default_rng(seed)
    c = np.repeat(rng.integers(0, 1 << 31, (n + 99) // 100, dtype=np.uint64), 100)[:n]
    out = np.zeros((n, 2), dtype=np.uint64)
    out[:, 0] = c + rng.integers(0, 1 << 31, n, dtype=np.uint64)
    return out


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "libhobbit_ref.so")
    if not os.path.exists(so):
        return None
    L = ctypes.CDLL(so)
    L.ref_init()
    L.ref_commit_standard.restype = ctypes.c_double
    L.ref_expander_init_store.restype = ctypes.c_longlong
    return L


def ref_commit_seconds(L, poly, k, trs):
    """Times the reference's commit_standard (Our_PC.cpp:146-171) on `poly` with k chunks; returns seconds."""
    B = len(poly) // k
    levels = np.zeros((2 * B - 1, 32), dtype=np.uint8)
    t = L.ref_commit_standard(poly.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(poly)), k, trs, 1,
                              levels.ctypes.data_as(ctypes.c_void_p), None)
    L.ref_release()
    return t, levels


def cpu_baseline(chunks, B, trs, seed=7):
    """The reference on a bounded sample: `chunks` chunks of the workload's chunk size (2^21 coefficients, trs 1024)."""
    L = load_ref()
    TRS = trs
    if chunks <= 0:
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "skipped (--cpu-chunks 0)"}
    if L is None:
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "oracle/_ref not built"}
    ctypes.CDLL(None).srand(1)
    L.ref_expander_init_store(ctypes.c_longlong(TRS))
    poly = make_poly(chunks * B, seed)
    t, _ = ref_commit_seconds(L, poly, chunks, TRS)
    return {"value": chunks * B / t, "unit": UNIT, "cores": 1, "kind": "reference", "seconds": t,
            "sample": "%d of %d chunks of the workload (2^%d coefficients, trs=%d, K=%d), unmodified reference commit_standard, "
                      "1 thread (the reference build has no OpenMP)" % (chunks, K, int(np.log2(chunks * B)), TRS, chunks)}


def run_reference(args, rank):
    """--impl reference: the reference's CPU commit_standard; each step = one chunk of the workload (2^21 coefficients)."""
    if rank != 0:
        return
    L = load_ref()
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (F_p^2, p=2^61-1) + u32 BLAKE3",
            "data": "synthetic"}
    if L is None:
        print(json.dumps(dict(base, unavailable="oracle/_ref/libhobbit_ref.so was not prebuilt (needs /root/reference at build time)")))
        return
    B = (1 << LOGN) // K
    ctypes.CDLL(None).srand(1)
    L.ref_expander_init_store(ctypes.c_longlong(TRS))
    poly = make_poly(B, 11)
    for _ in range(args.warmup):
        ref_commit_seconds(L, poly, 1, TRS)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_commit_seconds(L, poly, 1, TRS)
    dt = time.perf_counter() - t0
    v = args.steps * B / dt
    sample = "each step = 1 of the workload's %d chunks (2^21 coefficients, trs=%d) through the unmodified reference commit_standard, 1 thread" % (K, TRS)
    print(json.dumps(dict(base, value=v, ms_per_step=1e3 * dt / args.steps,
                          config={"workload": "Our_PC commit_standard N=2^26 K=32 trs=1024 linear_time (Orion columns) + BLAKE3 Merkle", "sample": sample},
                          cpu_baseline={"value": v, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample},
                          e2e={"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, gpu_launches=0)))


def host_mirror(local):
    """libhobbit_host.so: the C++ mirror of the reference API; owns the libc-RNG-ordered generators (expander graphs) and ONE context."""
    H = ctypes.CDLL(os.path.join(ROOT, "hobbit_b200", "libhobbit_host.so"))
    H.hobbit_c_backend.restype = ctypes.c_void_p
    H.hobbit_c_expander_init_store.restype = ctypes.c_longlong
    return H, H.hobbit_c_backend(int(local))


def run_tool(cmd, timeout=150):
    """Runs one of the reference-free C++ tools and returns its JSON line (rank 0 prints it), or {"error": ...}."""
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if p.returncode or not lines:
            return {"error": "rc=%d %s" % (p.returncode, (p.stdout + p.stderr)[-300:])}
        return json.loads(lines[-1])
    except Exception as e:                                   # noqa: BLE001
        return {"error": repr(e)}


def extras(world, rank):
    """The other BASELINE configurations through the host mirror (C++), outside every timed region of this script."""
    pc, mlp = os.path.join(ROOT, "hobbit_b200", "pc_prove"), os.path.join(ROOT, "hobbit_b200", "mlp_prove")
    env1 = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK=os.environ.get("LOCAL_RANK", "0"))
    out = {}
    if world == 1:
        def one(cmd):
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env1)
            lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
            return json.loads(lines[-1]) if lines and not p.returncode else {"error": "rc=%d %s" % (p.returncode, (p.stdout + p.stderr)[-300:])}
        out["config1_test_PC_2e20_K32"] = one([pc, "pc", "20", "32", "1"])
        # --async-levels: the Merkle levels of every commitment reach the caller's host vectors in the background, under the provers (they
        # are complete before open() frees them: the reference's own commit -> prove -> open sequence)
        out["config3_MLP_pigeon_9_18_18_1_4_1024_256_256_16"] = one([mlp, "18", "1024", "256", "256", "16", "--reps", "3", "--async-levels"])
        out["config4_AES_pigeon_5_19_8_1"] = one([mlp, "19", "aes", "8", "--reps", "3", "--async-levels"])
        out["config4_SQL_pigeon_6_19_17_1"] = one([mlp, "19", "sql", "17", "--reps", "3", "--async-levels"])
        out["config5_test_Elastic_PC_2e28_pinned"] = one([pc, "elastic", "28", "20", "2", "--pinned", "--resident-levels", "--reps", "2"])
        out["config5_test_Elastic_PC_2e28_hbm_chunk"] = one([pc, "elastic", "28", "20", "2", "--resident-levels", "--reps", "2"])
    else:
        # every rank of this job starts the same tool with its own RANK / LOCAL_RANK: the tools rendezvous among themselves (MASTER_PORT + 29)
        logn = {2: 28, 4: 29, 8: 30}.get(world, 28)
        # prover-only form (--resident-levels): the big Merkle levels stay in HBM on every rank (open() only needs their sizes); the default
        # form copies all 256 MiB of levels into every rank's MT_hashes, which at 8 ranks on one host is most of the commit time
        r = run_tool([pc, "elastic", str(logn), "20", "2", "--pinned", "--resident-levels", "--reps", "2"])
        if rank == 0:
            out["config5_test_Elastic_PC_2e%d_pinned_%dgpus" % (logn, world)] = r
        r = run_tool([pc, "elastic", str(logn), "20", "2", "--resident-levels", "--reps", "2"])
        if rank == 0:
            out["config5_test_Elastic_PC_2e%d_hbm_chunk_%dgpus" % (logn, world)] = r
        r = run_tool([mlp, "19", "sql", "17", "--reps", "2", "--async-levels"])
        if rank == 0:
            out["config4_SQL_pigeon_6_19_17_1_%dgpus" % world] = r
        r = run_tool([mlp, "19", "aes", "8", "--reps", "2", "--async-levels"])
        if rank == 0:
            out["config4_AES_pigeon_5_19_8_1_%dgpus" % world] = r
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hobbit_b200")
    ap.add_argument("--logn", type=int, default=LOGN, help="debug only: smaller polynomial (the reported workload is 2^26)")
    ap.add_argument("--cpu-chunks", type=int, default=4, help="chunks of the workload the CPU baseline is timed on")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations (C++ tools)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import hobbit_b200

    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    # one context, owned by the host mirror: the expander graphs are drawn by the product's own generator (hobbit::expander_init_store,
    # libc rand()/random() in the reference's call order) — nothing under oracle/ is loaded by this arm before the CPU baseline
    H, handle = host_mirror(local)
    ctx = hobbit_b200.Context.from_handle(handle)
    N = 1 << args.logn
    B = N // K
    trs = TRS if args.logn == LOGN else max(16, N // (K << 11))
    ctypes.CDLL(None).srand(1)
    cw = int(H.hobbit_c_expander_init_store(ctypes.c_longlong(trs)))

    def gen_poly(r):
        """generate_randomness-shaped synthetic input (utils.cpp:873-883: real-only values c + rand()), drawn on the device"""
        g = torch.Generator(device="cuda"); g.manual_seed(1234 + r)
        t = torch.zeros((N, 2), dtype=torch.int64, device="cuda")
        t[:, 0] = torch.randint(0, 1 << 32, (N,), dtype=torch.int64, device="cuda", generator=g)
        return t

    # N > 1: ONE commitment of N*world coefficients (K*world chunks of the same size) sharded over the ranks: rank g owns chunks
    # [g*K, (g+1)*K) and the leaf range [g*B/world, (g+1)*B/world); N == 1: the plain C-ABI call.
    poly_dev = gen_poly(rank)
    poly_host = ctx.pinned((N, 2), np.uint64)
    torch.from_numpy(poly_host.view(np.int64)).copy_(poly_dev)
    levels_host = ctx.pinned((2 * B - 1, 32), np.uint8)
    levels_dev = torch.empty(((2 * B - 1) * 32,), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream())

    if world > 1:
        ctx.dist_init_torch(K * world * (B // world) * 32 + 2 * B * 32 + (1 << 20) + 3 * 16 * (1 << 22) * 0)

    def step_resident():
        if world > 1:
            ctx.dist_commit_standard(poly_dev.data_ptr(), K * world, B, trs, 1, levels_out=levels_dev.data_ptr())
        else:
            ctx.commit_standard(poly_dev.data_ptr(), K, trs, 1, levels_out=levels_dev.data_ptr(), N=N)

    def step_e2e():
        # host polynomial in (pinned), every Merkle level out to the host — on rank 0 (the caller that holds the commitment); the other
        # ranks keep their copy of the tree in HBM
        if world > 1:
            ctx.dist_commit_standard(poly_host, K * world, B, trs, 1, levels_out=levels_host if rank == 0 else levels_dev.data_ptr())
        else:
            ctx.commit_standard(poly_host, K, trs, 1, levels_out=levels_host)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launch_count() - l0

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_resident()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 5.0:     # nvidia-smi needs a moment to start streaming
        time.sleep(0.05)
    sampler.rows.clear()
    ms, launches = timed(step_resident, args.steps)            # headline: no per-launch events inside this region
    clocks = sampler.finish()
    for _ in range(max(1, args.warmup - 1)):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    # separate pass: CUDA events around every launch (on the library's stream) -> per-kernel launch times for the roofline
    ctx.profile(True)
    psteps = max(2, min(5, args.steps))
    for _ in range(psteps):
        step_resident()
    prof = ctx.profile_report()
    ctx.profile(False)

    value = world * args.steps * N / (ms * 1e-3)
    e2e_v = world * args.steps * N / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel -------------------------------------------------------------------------------------------
    prof = {k: v for k, v in prof.items() if not k.startswith("dist_barrier")}
    name, rec = max(prof.items(), key=lambda kv: kv[1]["total_ms"])
    per_launch_ms = rec["total_ms"] / rec["launches"]
    coeffs_per_launch = float(N) * psteps / rec["launches"]
    # algorithmic bytes per coefficient: encode = read the RS-encoded rows (32 B) + write the parity rows / zero tail (32 B)
    # + write the inner leaf digest (32 B); NTT = 16 B in + 32 B out; chain = 32 B of inner digests
    alg_bytes = {"encode_cols_kernel": 96.0, "ntt_tile_kernel": 48.0, "ntt_tile_lazy_kernel": 48.0, "md_chain_kernel": 32.0}.get(name, 64.0) * coeffs_per_launch
    peak, how = peaks()
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
    counts = {}
    cp = os.path.join(ROOT, "profiles", "kernel_counts_r02.json")   # per-kernel ncu counters of this command (instructions, DRAM bytes), by kernel name
    if os.path.exists(cp):
        counts = json.load(open(cp)).get(name, {})
    traffic = counts.get("dram_bytes_per_coefficient")
    total_ms = sum(x["total_ms"] for x in prof.values())
    roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic * coeffs_per_launch if traffic else None, "peak_source": how, "avg_launch_ms": per_launch_ms,
                "launches_timed": rec["launches"], "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_time_share": {k: round(v["total_ms"] / total_ms, 4) for k, v in prof.items()},
                "timing": "CUDA events around every launch on the library's stream, in a pass of %d steps separate from the headline region" % psteps}
    # The BINDING roof of this kernel is not HBM: 61-bit modular multiply-adds and BLAKE3 are integer work and sm_100a has no 64-bit
    # multiplier.  Reported next to the HBM numbers: (1) the issue-slot roof — the kernel's warp-instruction count (ncu counters of THIS
    # kernel, by name, from the committed launch list) over the live launch time, against the 4 warp-instructions per clock and SM the
    # four schedulers can issue; (2) the same for the ALU pipe and the FMA-heavy pipe; (3) what the two integer pipes deliver on THIS GPU
    # in isolation, measured now (hb_ubench_pipes): IMAD.WIDE.U32 alone, SHF/LOP3 alone, a 1:3 mix.
    try:
        pipes = ctx.ubench_pipes()
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        per = (clocks.get("sm_mhz") or 0) * 1e6 * sms                                     # SM-cycles per second at the sampled clock
        if per:
            pipes["per_clk_per_sm_at_sampled_clock"] = {k: pipes[k] / per for k in ("imad_wide", "alu", "mix_1_wide_3_alu")}
        roofline["measured_pipe_rates"] = pipes
        if counts.get("warp_inst_per_coefficient") and per:
            cyc = per_launch_ms * 1e-3 * per                                              # SM-cycles of one launch
            rate = lambda key: counts[key] * coeffs_per_launch / cyc if counts.get(key) else None
            issue = rate("warp_inst_per_coefficient")
            roofline["binding"] = {"resource": "instruction issue (integer pipes: IMAD / IMAD.WIDE on FMA-heavy, LOP3 / SHF / IADD3 on ALU)",
                                   "unit": "warp-inst/clk/SM", "achieved": issue, "peak": 4.0, "frac": issue / 4.0,
                                   "alu_pipe": {"achieved": rate("alu_inst_per_coefficient"), "measured_peak": pipes["per_clk_per_sm_at_sampled_clock"]["alu"]},
                                   "fmaheavy_pipe": {"achieved": rate("fmaheavy_inst_per_coefficient"),
                                                     "note": "IMAD.WIDE alone sustains %.2f on this GPU; plain IMAD issues at twice that" % pipes["per_clk_per_sm_at_sampled_clock"]["imad_wide"]},
                                   "source": "instruction counts: profiles/kernel_counts_r02.json (ncu smsp__inst_executed.sum / sm__inst_executed_pipe_*.sum of this "
                                             "kernel's full-size launches, by name); time and clock: this run"}
    except Exception as e:                                   # noqa: BLE001
        roofline["measured_pipe_rates"] = {"error": repr(e)}

    # ---- N > 1: sharded vs single-GPU, outside the timed region --------------------------------------------------------------------
    parity = None
    if world > 1:
        parity = {}
        ctx.dist_commit_standard(poly_dev.data_ptr(), K * world, B, trs, 1, levels_out=levels_dev.data_ptr())
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        if rank == 0:
            # rank 0 regenerates every rank's polynomial (same device generator, same seeds) and commits the whole thing on ONE GPU
            whole = torch.cat([poly_dev] + [gen_poly(r) for r in range(1, world)], dim=0)
            single = torch.empty(((2 * B - 1) * 32,), dtype=torch.uint8, device="cuda")
            ctx.commit_standard(whole.data_ptr(), K * world, trs, 1, levels_out=single.data_ptr(), N=N * world)
            ok[0] = int(torch.equal(single, levels_dev))
            del whole, single
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity["commit"] = bool(ok.item())
        parity["commit_what"] = "every Merkle level (2^%d - 1 digests) of the sharded 2^%d-coefficient commitment == hb_commit_standard of the same polynomial on one GPU" % (args.logn - 4, args.logn + int(np.log2(world)))
        # one sharded sumcheck proof (3-product, 2^22-entry tables, identical on every rank) against the single-GPU proof
        g = torch.Generator(device="cuda"); g.manual_seed(99)
        tabs = [torch.randint(0, (1 << 61) - 1, (1 << 22, 2), dtype=torch.int64, device="cuda", generator=g) for _ in range(3)]
        pr = np.array([[5, 7]], dtype=np.uint64)
        dv = [hobbit_b200.DevF.from_torch(t) for t in tabs]
        ctx.dist_shard(False)
        want, _ = ctx.sumcheck3(dv[0], dv[1], dv[2], pr)
        ctx.dist_shard(True)
        r0 = ctx.dist_stats()["reductions"]
        got, _ = ctx.sumcheck3(dv[0], dv[1], dv[2], pr)
        ctx.dist_shard(False)
        ok[0] = int(np.array_equal(want, got) and ctx.dist_stats()["reductions"] > r0)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity["sumcheck"] = bool(ok.item())
        parity["sumcheck_what"] = "_generate_3product_sumcheck_proof over 2^22-entry tables sharded by hypercube prefix == the single-GPU proof, all 114 field elements"
        del tabs

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u64 (F_p^2, p=2^61-1) + u32 BLAKE3", "data": "synthetic",
           "config": {"workload": "Our_PC commit_standard N=2^%d K=%d trs=%d linear_time (RS rows + Orion expander columns) + BLAKE3 Merkle (test_PC option 4 commit)" % (args.logn, K, trs),
                      "codeword_len": cw, "l2": "inputs larger than L2 (1 GiB polynomial, 4 GiB tensor per step)",
                      "multi_gpu": ("ONE commitment of 2^%d x %d coefficients behind the C ABI (hb_dist_commit_standard): chunks sharded by rank, inner leaf digests stored by the "
                                    "encode kernel into the owner rank's window over NVLink, subtrees scattered to every rank" % (args.logn, world)) if world > 1 else "single GPU"},
           "e2e": {"value": e2e_v, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": int(poly_host.nbytes),
                   "d2h_bytes_per_step": int(levels_host.nbytes),
                   "note": "pinned host polynomial in on every rank, every Merkle level out on rank 0" if world > 1 else "pinned host polynomial in, every Merkle level out"},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if parity is not None:
        out["parity_check"] = parity
    # free the big buffers before the C++ tools run on the same GPUs
    del poly_dev, levels_dev
    torch.cuda.empty_cache()
    if world > 1:
        ctx.dist_disconnect()
    if not args.no_extras:
        if world > 1:
            dist.barrier()
        ex = extras(world, rank)
        if rank == 0:
            out["extras"] = ex
    if rank == 0:
        if world == 1:
            out["cpu_baseline"] = cpu_baseline(args.cpu_chunks, B, trs)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
