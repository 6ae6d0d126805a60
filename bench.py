#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native G1 hot path.

Workload (BASELINE.json configs[1]): Whisk shape, shuffled_elements = 124
(n = 128): GenerateWhiskShuffleProof + IsValidWhiskShuffleProof round trips.
One *step* = one batch of B independent round trips (synthetic trackers),
driven through the C ABI (cdl_whisk_*_batch) with HOST buffers.

  value      proofs/s counting only device time: B*K / device busy time, the
             union of the CUDA-event intervals of every kernel the step
             launched on any lane's stream (inputs resident in HBM; the host's
             Fiat-Shamir work between launches excluded)
  e2e.value  proofs/s wall clock through the C ABI, host<->device copies and the
             host-side transcript work inside the timed region
  roofline   the dominant kernel class against the integer-pipe peak
             (IMAD.WIDE issue rate measured live / 300 multiply-adds per
             381-bit Montgomery product), HBM GB/s secondary
  cpu_baseline  the CPU oracle ("port": C 6x64 Montgomery restatement driven by
             the Python protocol restatement) on the host cores

  msm        the standalone G1 MSM sweep (config 5) at a few sizes: device-
             resident vectors, CUDA-event time of the Pippenger kernel chain;
             at N > 1 GPUs the window-partitioned form (NCCL all-gather of
             one partial sum per rank inside libcurdle_b200.so)

`--impl reference` times that CPU path alone on all host cores.
Multi-GPU: one process per GPU (torchrun), independent proofs sharded across
ranks, no data-path collective (weak scaling); NCCL only for the barrier and the
max-over-ranks time.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ELL = 124
METRIC = "whisk_shuffle_proof_roundtrips_per_s (GenerateWhiskShuffleProof + IsValidWhiskShuffleProof, n=128)"
UNIT = "proofs/s"
MODMUL_IMADS = 300.0  # 2*12^2 + 12 wide multiply-adds per 381-bit Montgomery product


# ----------------------------------------------------------------------------- reference arm (CPU oracle)
def _ref_worker(args):
    """One worker: `count` Whisk round trips on the CPU oracle, returns seconds."""
    seed, count = args
    sys.path.insert(0, ROOT)
    from oracle import protocol as P, whisk as W
    from oracle.cbackend import CBackend
    from oracle.rand import Rand

    cb = CBackend(threads=1)
    P.set_backend(cb)
    rand = Rand(0, backend=cb)
    crs = P.generate_crs(ELL, rand)
    pre = W.generate_shuffle_trackers(Rand(1000 + seed, backend=cb), ELL)
    t0 = time.perf_counter()
    for i in range(count):
        r = Rand(3000 + seed * 1000 + i, backend=cb)
        post, proof = W.generate_whisk_shuffle_proof(crs, pre, r)
        ok = W.is_valid_whisk_shuffle_proof(crs, pre, post, proof, r)
        assert ok
    return time.perf_counter() - t0


def run_reference_sample(workers: int, per_worker: int):
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        pool.map(_ref_worker, [(w, per_worker) for w in range(workers)])
    dt = time.perf_counter() - t0
    return workers * per_worker, dt


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, ROOT)
    from oracle.cbackend import build as build_oracle

    build_oracle()
    cores = os.cpu_count() or 1
    workers = cores
    per_worker = 2
    _ref_worker((999, 1))  # warm import / build, untimed
    for _ in range(args.warmup):
        run_reference_sample(workers, per_worker)
    t = []
    proofs = 0
    for _ in range(args.steps):
        n, dt = run_reference_sample(workers, per_worker)
        proofs += n
        t.append(dt)
    total = sum(t)
    value = proofs / total
    sample = (f"{workers} worker processes x {per_worker} Whisk n=128 round trip(s) per step, pool start-up "
              f"included; {args.steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (6x64-bit Montgomery limbs)",
        "data": "synthetic", "config": {"workload": "whisk_n128_roundtrip", "shuffled_elements": ELL,
                                         "proofs_per_step": workers * per_worker},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[1]) for s in self.samples if len(s) > 2 and s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for k, nm in enumerate(names):
                if len(s) > 5 + k and s[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def make_trackers(ctx, pkg, ell: int, seeds):
    """Synthetic pre-shuffle trackers (whisk_test.go generateShuffleTrackers): (r*G, k*r*G) compressed."""
    enc = importlib.import_module("go-curdleproofs_b200.encoding")
    gen = enc.aff_enc(enc.G1_GEN)
    out = []
    for s in seeds:
        r = pkg.Rand(s)
        ks, rs = [], []
        for _ in range(ell):
            ks.append(r.get_fr())
            rs.append(r.get_fr())
        rG = ctx.g1_scalar_mul_affine(gen * ell, b"".join(rs), broadcast=False)
        krG = ctx.g1_scalar_mul_affine(rG, b"".join(ks), broadcast=False)
        e1, e2 = ctx.g1_compress(rG), ctx.g1_compress(krG)
        out.append(b"".join(e1[48 * j:48 * j + 48] + e2[48 * j:48 * j + 48] for j in range(ell)))
    return out


def msm_alg_modmul(n: int) -> int:
    """min_c [6 n W + 28 2^(c-1) W + 9 c (W-1) + 14 W], W = ceil(256 / c)."""
    best = None
    for c in range(1, 25):
        W = -(-256 // c)
        v = 6 * n * W + 28 * (1 << (c - 1)) * W + 9 * c * (W - 1) + 14 * W
        best = v if best is None or v < best else best
    return best


def msm_sweep(ctx, pkg, log2_sizes, reps, world, rank, peak_modmul, barrier):
    """Config 5: points P_i = a_i*G (a_i = Rand(5).GetFr), scalars Rand(6).GetFr, both replicated on
    every rank; the result is checked against (sum a_i*s_i)*G computed with host Fr arithmetic
    through a 1-point scalar multiplication on the GPU."""
    enc = importlib.import_module("go-curdleproofs_b200.encoding")
    R, RR_INV, aff_enc, fr_enc = enc.R, enc.RR_INV, enc.aff_enc, enc.fr_enc

    nmax = 1 << max(log2_sizes)
    r5, r6 = pkg.Rand(5), pkg.Rand(6)
    dp, da, ds = ctx.dev_buffer(96 * nmax), ctx.dev_buffer(32 * nmax), ctx.dev_buffer(32 * nmax)
    gen = aff_enc(enc.G1_GEN)
    chunk = 1 << min(16, min(log2_sizes))
    acc = {lg: 0 for lg in log2_sizes}
    done = 0
    for o in range(0, nmax, chunk):
        m = min(chunk, nmax - o)
        A, S = r5.get_frs(m), r6.get_frs(m)
        dp.upload(gen * m, 96 * o)
        da.upload(A, 32 * o)
        ds.upload(S, 32 * o)
        part = 0
        if rank == 0:  # host-side check value: sum a_i * s_i (Montgomery decode folded into one factor)
            for i in range(m):
                part += int.from_bytes(A[32 * i:32 * i + 32], "little") * int.from_bytes(S[32 * i:32 * i + 32], "little")
        done += m
        for lg in log2_sizes:
            if done <= (1 << lg):
                acc[lg] += part
    ctx.g1_scalar_mul_affine_device(dp, da, nmax, False, dp)
    out = []
    for lg in log2_sizes:
        n = 1 << lg
        best, res = None, None
        for _ in range(reps + 1):
            barrier()
            res, ms = ctx.g1_msm_sharded_device(dp, ds, n)
            best = ms if best is None or ms < best else best
        want_k = acc[lg] * RR_INV * RR_INV % R
        want = ctx.g1_scalar_mul_affine(gen, fr_enc(want_k), broadcast=True)
        ok = res[:96] == want and res[96:144] != bytes(48)
        mm = msm_alg_modmul(n)
        out.append({"log2n": lg, "ms": best, "mpoints_per_s": n / best / 1e3, "gmodmul_per_s": mm / best / 1e6,
                    "frac_of_int_peak": mm / (best * 1e-3) / (peak_modmul * world), "check": "ok" if ok else "MISMATCH",
                    "windows_per_rank": pkg.comm_partition(n, world, rank)[3]})
    cpu = None
    if rank == 0 and not os.environ.get("CDL_BENCH_NO_CPU_MSM"):
        # the same MSM on the host cores with the CPU oracle port (bucket method, windows spread over
        # threads; the reference's gnark-crypto MultiExp cannot be built here), bounded to 2^16 terms
        try:
            from oracle.cbackend import CBackend  # CPU baseline leg: the one place the GPU arm runs the oracle

            RP_INV, FP_P = enc.RP_INV, enc.P

            lg = min(16, min(log2_sizes))
            n = 1 << lg
            raw_p = dp.download(96 * n)
            raw_s = ds.download(32 * n)
            cp = b"".join((int.from_bytes(raw_p[48 * i:48 * i + 48], "little") * RP_INV % FP_P).to_bytes(48, "little")
                          for i in range(2 * n))
            cs = b"".join((int.from_bytes(raw_s[32 * i:32 * i + 32], "little") * RR_INV % R).to_bytes(32, "little")
                          for i in range(n))
            cb = CBackend(accelerate_keccak=False)
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            res_cpu = cb.msm_raw(cp, cs, n, cores)
            dt = time.perf_counter() - t0
            gpu_res, _ = ctx.g1_msm_device(dp, ds, n)  # local (non-collective): only rank 0 is here
            same = all((int.from_bytes(gpu_res[48 * k:48 * k + 48], "little") * RP_INV % FP_P) ==
                       int.from_bytes(res_cpu[48 * k:48 * k + 48], "little") for k in range(2))
            cpu = {"log2n": lg, "ms": dt * 1e3, "mpoints_per_s": n / dt / 1e6, "cores": cores, "kind": "port",
                   "check": "equal to the GPU result" if same else "MISMATCH"}
        except Exception as e:  # reported baseline, never required
            cpu = {"failed": str(e)}
    for d in (dp, da, ds):
        d.close()
    return out, cpu


def single_proof_latency(ctx, pkg, ells=(60, 124, 508)):
    """BASELINE.json configs 1-3: curdleproof.Prove / Verify for one proof at a time through the C ABI
    (setup of curdleproof_test.go:239-274 with perm = Rand(42).GeneratePermutation), best of 3, wall clock."""
    out = []
    for ell in ells:
        r = pkg.Rand(0)
        crs = ctx.generate_crs(ell, r)
        k = r.get_fr()
        Rs = ctx.rand_get_g1_affines(r, ell)
        Ss = ctx.rand_get_g1_affines(r, ell)
        perm = pkg.Rand(42).generate_permutation(ell)
        Ts, Us, M, rs_m = ctx.shuffle_permute_commit(crs, Rs, Ss, perm, k, r)
        bp = bv = None
        for it in range(4):
            t0 = time.perf_counter()
            proof = ctx.prove(crs, Rs, Ss, Ts, Us, M, perm, k, rs_m, pkg.Rand(42))
            t1 = time.perf_counter()
            ok = ctx.verify(crs, proof, Rs, Ss, Ts, Us, M, pkg.Rand(43))
            t2 = time.perf_counter()
            assert ok
            if it:
                bp = t1 - t0 if bp is None else min(bp, t1 - t0)
                bv = t2 - t1 if bv is None else min(bv, t2 - t1)
        out.append({"shuffled_elements": ell, "prove_ms": bp * 1e3, "verify_ms": bv * 1e3, "proof_bytes": len(proof)})
        crs.close()
    return out


def gpu_main(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    # ranks of one box share its host cores (Fiat-Shamir work between launches)
    os.environ.setdefault("CDL_HOST_THREADS", str(max(2, (os.cpu_count() or 1) // max(1, world))))
    pkg = importlib.import_module("go-curdleproofs_b200")
    ctx = pkg.Context(local)  # raises without a GPU / without the built library: no fallback
    if args.lanes:
        ctx.set_lanes(args.lanes)
    if world > 1:  # the library's own NCCL communicator (window-partitioned MSM)
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            uid = torch.tensor(list(pkg.comm_unique_id()), dtype=torch.uint8, device=f"cuda:{local}")
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().tolist()), rank, world)
    info = ctx.device_info()
    B = args.batch
    crs = ctx.generate_crs(ELL, pkg.Rand(0))
    # a few distinct tracker sets, cycled over the batch (every instance still gets its own RNG stream)
    base_sets = make_trackers(ctx, pkg, ELL, [1000 + rank * 7919 + i for i in range(min(B, 8))])
    pre = b"".join(base_sets[i % len(base_sets)] for i in range(B))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2

    # integer-pipe peak, measured live (burst): IMAD.WIDE/s over the whole chip
    imad_wide_per_s, _ = ctx.int_peak(1, 4000)
    peak_modmul = imad_wide_per_s / MODMUL_IMADS

    step_no = [0]

    def step():
        flush.zero_()  # L2 flush between iterations
        s = step_no[0]
        step_no[0] += 1
        rands = [pkg.Rand((rank << 40) | (s << 20) | i) for i in range(B)]
        post, proofs, status = ctx.whisk_generate_shuffle_proof_batch(crs, pre, rands)
        ok, st = ctx.whisk_is_valid_shuffle_proof_batch(crs, pre, post, proofs, rands)
        return status, ok, st

    for _ in range(args.warmup):
        status, ok, st = step()
        assert status == [0] * B and ok == [1] * B and st == [0] * B, "warm-up round trip failed"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    ctx.engine_stats(reset=True)
    l0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        status, ok, st = step()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    assert status == [0] * B and ok == [1] * B and st == [0] * B, "timed round trip failed"
    stats_timed = ctx.engine_stats()
    launches = ctx.launch_count() - l0
    busy_ms = ctx.engine_busy_ms()
    # roofline pass: the same step with ONE lane, so that the CUDA-event interval of a kernel is
    # its own duration (with several lanes the streams share the GPU and every interval also
    # contains the other lanes' kernels); one untimed warm-up, then one measured step
    ctx.set_lanes(1)
    step()
    ctx.engine_stats(reset=True)
    status, ok, st = step()
    torch.cuda.synchronize()
    assert status == [0] * B and ok == [1] * B and st == [0] * B, "roofline-pass round trip failed"
    stats = ctx.engine_stats()
    ctx.set_lanes(args.lanes or 4)
    dev_ms = sum(v["ms"] for v in stats.values())
    # config 4 of BASELINE.json: batched verification only, every 8th instance mutated (pre / post
    # trackers swapped) so that reject paths are exercised; wall clock through the C ABI
    tb = ELL * 96
    post, proofs, status = ctx.whisk_generate_shuffle_proof_batch(crs, pre, [pkg.Rand((rank << 40) | (77 << 20) | i) for i in range(B)])
    pre_m, post_m, proofs_m = bytearray(pre), bytearray(post), bytearray(proofs)
    want_ok, want_st = [1] * B, [0] * B
    for i in range(0, B, 8):  # every 8th instance is mutated, cycling through three kinds
        kind = (i // 8) % 3
        want_ok[i] = 0
        if kind == 0:    # pre and post trackers swapped
            pre_m[i * tb:(i + 1) * tb], post_m[i * tb:(i + 1) * tb] = post[i * tb:(i + 1) * tb], pre[i * tb:(i + 1) * tb]
        elif kind == 1:  # one bit of the proof's last scalar flipped (bytes 4504..4535 of the 4576)
            proofs_m[i * 4576 + 4535] ^= 1
        else:            # first post-tracker point replaced by the infinity encoding: "randomizer is zero"
            post_m[i * tb:i * tb + 48] = bytes([0xC0]) + bytes(47)
            want_st[i] = -5
    vt = []
    verdict_ok = True
    for it in range(3):  # first pass is the warm-up
        barrier()
        t1 = time.perf_counter()
        vok, vst = ctx.whisk_is_valid_shuffle_proof_batch(crs, pre_m, post_m, proofs_m,
                                                          [pkg.Rand((rank << 40) | (78 << 20) | i) for i in range(B)])
        torch.cuda.synchronize()
        vt.append(time.perf_counter() - t1)
        verdict_ok = verdict_ok and vok == want_ok and vst == want_st
    tv = torch.tensor([sum(vt[1:])], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    verify_only = {"workload": "batched IsValidWhiskShuffleProof, n=128, every 8th instance mutated (trackers swapped / "
                               "scalar bit flipped / infinity tracker, in turn)",
                   "value": world * B * 2 / float(tv[0]), "unit": "verifications/s", "batch_per_gpu": B,
                   "timing": "wall clock through the C ABI with host buffers, 2 batches after 1 warm-up",
                   "verdicts": "as expected" if verdict_ok else "MISMATCH"}
    msm, msm_cpu = None, None
    if not args.no_msm:
        msm, msm_cpu = msm_sweep(ctx, pkg, [16, 20, 22] if not args.msm_sizes else [int(x) for x in args.msm_sizes.split(",")],
                        2, world, rank, peak_modmul, barrier)

    t = torch.tensor([wall, busy_ms / 1e3], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_max, dev_max = float(t[0]), float(t[1])
    total_proofs = world * B * args.steps

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        top = max(stats, key=lambda k: stats[k]["ms"])
        tk = stats[top]
        n_launch = max(1, tk["launches"])
        avg_ms = tk["ms"] / n_launch
        achieved = tk["modmul"] / (tk["ms"] * 1e-3) if tk["ms"] else 0.0
        traffic = None
        try:  # dram bytes per launch of the dominant kernel class from an ncu capture of this command
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("batch") == B:
                traffic = tj.get(top)
        except Exception:
            pass
        timed = stats_timed[top]
        roofline = {
            "bound": "int_alu", "kernel": top + (" (k_msm_recode + k_msm_warp + k_msm_chunk_sum + k_msm_combine_tp)"
                                                 if top == "msm_small" else ""),
            "achieved": achieved / 1e9, "peak": peak_modmul / 1e9,
            "unit": "Gmodmul/s", "frac": achieved / peak_modmul if peak_modmul else None, "traffic": traffic,
            "measured_in": "one extra step of the same batch with 1 lane, CUDA events on the launching stream "
                           "(kernel intervals of concurrent lanes overlap and would count each other's time)",
            "timed_region": {"lanes": args.lanes or 4, "launches": timed["launches"],
                             "avg_launch_ms_overlapped": timed["ms"] / max(1, timed["launches"])},
            "carry_chain_ceiling": "IMAD.WIDE.U32 with a carry issues at half rate on sm_100a (profiles/"
                                   "r1_probe_instruction_rates.txt): a 32-bit-limb Montgomery product cannot exceed "
                                   "frac 0.5",
            "peak_source": "measured live: IMAD.WIDE.U32 issue rate (cdl_int_peak kind 1, burst) / 300 per modmul",
            "avg_launch_ms": avg_ms, "launches": tk["launches"],
            "algorithmic_modmul_per_launch": tk["modmul"] / n_launch,
            "hbm": {"achieved_gbs": tk["bytes"] / (tk["ms"] * 1e-3) / 1e9 if tk["ms"] else 0.0, "peak_gbs": hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "per_kernel": {k: {"ms": v["ms"], "launches": v["launches"],
                               "gmodmul_per_s": (v["modmul"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] else 0.0,
                               "share_of_device_time": v["ms"] / dev_ms if dev_ms else 0.0} for k, v in stats.items()},
        }
        h2d = B * (ELL * 96) + B * (2 * ELL * 96 + 4576)   # generate: pre ; validate: pre + post + proof
        d2h = B * (ELL * 96 + 4576) + B * 8                # generate: post + proof ; validate: verdict + status
        line = {
            "metric": METRIC, "value": total_proofs / dev_max, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 (12x32-bit Montgomery limbs, IMAD.WIDE)",
            "data": "synthetic",
            "config": {"workload": "whisk_n128_roundtrip", "shuffled_elements": ELL, "proofs_per_gpu_per_step": B,
                       "parallelism": f"proof-parallel x{world}, no data-path collective",
                       "l2": "flushed between steps (256 MiB write); the step's own working set (about 120 kB of points per "
                             "instance) exceeds the 126 MB L2 as well",
                       "value_definition": "proofs / device busy time = union of CUDA-event kernel intervals over all lanes "
                                           "(host Fiat-Shamir excluded)",
                       "lanes": args.lanes or 4, "host_threads_cap": int(os.environ.get("CDL_HOST_THREADS", "0")),
                       "device": info["name"], "sm_count": info["sm_count"]},
            "e2e": {"value": total_proofs / wall_max, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "note": "C ABI with host buffers; stage descriptors/scalars (about 36 B per MSM term) are extra H2D"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "clocks": sampler.summary(),
        }
        line["verify_only"] = verify_only
        if world == 1 and not args.no_latency:
            line["single_proof"] = {"what": "curdleproof.Prove / Verify, one proof at a time (configs 1-3), wall clock",
                                    "reference_published_ms": {"prove": [96.4, 150.2, 412.5], "verify": [12.0, 12.3, 20.8],
                                                               "hardware": "Ryzen 7 3800XT, README.md:19-26"},
                                    "sizes": single_proof_latency(ctx, pkg)}
        if msm is not None:
            line["msm"] = {"workload": "standalone G1 MSM, random points/scalars, device resident"
                                       + (f", windows dealt to {world} ranks + NCCL all-gather" if world > 1 else ""),
                           "sizes": msm, "cpu_baseline": msm_cpu}
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle.cbackend import build as build_oracle

                build_oracle()
                cores = os.cpu_count() or 1
                workers = min(cores, 32)
                n, dt = run_reference_sample(workers, 1)
                line["cpu_baseline"] = {
                    "value": n / dt, "unit": UNIT, "cores": workers, "kind": "port",
                    "sample": f"{workers} processes x 1 Whisk n=128 round trip on the CPU oracle (C 6x64 Montgomery "
                              f"port + Python orchestration), {dt:.1f} s wall incl. process start-up",
                    "published_reference": "README.md:20,24 — Prove 150.2 ms + Verify 12.3 ms on a Ryzen 7 3800XT "
                                           "(16 threads) = 6.2 proofs/s excl. ShufflePermuteCommit and decompression",
                }
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=2048, help="independent Whisk round trips per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=8, help="concurrent sub-batches per GPU (0 = library default)")
    ap.add_argument("--no-msm", action="store_true", help="skip the standalone MSM sweep")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-proof latency lines")
    ap.add_argument("--msm-sizes", default="", help="comma separated log2 sizes for the MSM sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_main(args)
    else:
        gpu_main(args)


if __name__ == "__main__":
    main()
