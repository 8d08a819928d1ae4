#!/usr/bin/env python
"""bench.py -- Zip commit throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--nv 24] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one Zip commit (RAA encode of every row + one BLAKE3 Merkle tree per row) of one synthetic
2^nv-evaluation Int<1> MLE per GPU ("weak" scaling: N GPUs commit N independent MLEs / row ranges, no
data-path collective -- SURVEY.md 8e).

  value     evaluations/s, whole job, inputs resident in HBM, CUDA-event timed on the launch stream
  e2e       the same metric through the C ABI with HOST buffers: pinned H2D of the evaluations and D2H of the
            roots inside the timed region (zipgpu_commit_resident: rows/layers stay in HBM for `open`)
  roofline  the encoder kernel (HBM-bound): 72 algorithmic bytes per evaluation / its mean launch duration
  hasher    the BLAKE3 kernels (INT32-ALU-bound): compressions/s, lane-instructions/s against the measured peak
  cpu_baseline  the oracle's multithreaded restatement of the reference's CPU commit on a bounded row sample

`--impl reference` times the CPU restatement only (rank 0), same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KECCAK_SEEDS = (0xB9736F582676E7E8, 0xD7397E6260CE9C3E)  # fresh KeccakTranscript, as zip_benches.rs:102-106
ENC_BYTES_PER_EVAL = 72.0          # 8 B read + 2 x 32 B written
HASH_INSTR_PER_COMPRESSION = 680.0  # 7 rounds x 8 G x 12 + 8, minimum INT32 lane-instructions
# alu-pipe (LOP3/SHF/PRMT/IADD3: 64 lanes/clk/SM on B200, scratch/pipes.cu) lane-instructions per compression: the
# irreducible 7 x 8 x (4 xor + 4 rotate) + 8 feed-forward xors = 456, and what the SASS of merkle_subtree_kernel
# actually issues (cuobjdump count, leaf + node average): 480
HASH_ALU_MIN_PER_COMPRESSION = 456.0
HASH_ALU_SASS_PER_COMPRESSION = 480.0
ALU_LANES_PER_CLK_PER_SM = 64.0


# stdout must carry exactly ONE JSON line: everything any library prints on fd 1 (NCCL's version banner, build chatter)
# is diverted to stderr, and the result line is written to the saved descriptor
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def shape_for(nv: int):
    """row_len / num_rows / cw for a 2^nv MLE: row_len = isqrt(2^nv).next_power_of_two() (code_raa.rs:42-43),
    num_rows = (2^nv / row_len).next_power_of_two() (structs.rs:82), cw = 2 * row_len (rep = 2)"""
    import math

    r = math.isqrt(1 << nv)
    row_len = 1 << (r - 1).bit_length() if r > 1 else 1
    q = (1 << nv) // row_len
    num_rows = 1 << (q - 1).bit_length() if q > 1 else 1
    return row_len, num_rows, 2 * row_len


def gen_evals(nv: int, rank: int) -> np.ndarray:
    """uniform 64-bit words, like DenseMultilinearExtension::rand (poly_z/mle/dense.rs:140-145)"""
    rng = np.random.Generator(np.random.PCG64(0x21C0 + nv + 1000 * rank))
    return rng.integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)


def workload_name(nv: int) -> str:
    row_len, num_rows, cw = shape_for(nv)
    return (f"zip_commit nv={nv}: 2^{nv} Int<1> evals as {num_rows} rows x {row_len}, RAA rep=2 -> cw={cw} Int<4>, "
            f"{num_rows} BLAKE3 Merkle trees of depth {cw.bit_length() - 1}; perm seeds = fresh KeccakTranscript")


# --------------------------------------------------------------------------------------------------
# clocks (NVML, sampled in a thread during the timed regions)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # no NVML: report nulls
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# CPU baseline (oracle restatement; the only place bench.py executes oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_commit_sample(nv: int, sample_rows: int, threads: int, faithful: bool, evals: np.ndarray, reps: int = 1):
    from oracle import cbind

    cbind.build()
    row_len, num_rows, cw = shape_for(nv)
    sample_rows = min(sample_rows, num_rows)
    p1, p2 = cbind.perm_from_seed(cw, KECCAK_SEEDS[0]), cbind.perm_from_seed(cw, KECCAK_SEEDS[1])
    ev = np.ascontiguousarray(evals[: sample_rows * row_len])
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        rc, _, _, roots = cbind.commit_mt(ev, sample_rows, row_len, 2, KECCAK_SEEDS[0], KECCAK_SEEDS[1], p1, p2,
                                          threads=threads, faithful=faithful, want_rows=True, want_layers=True)
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    return sample_rows * row_len / best, best, sample_rows, roots


def run_reference_arm(args, rank: int, world: int):
    """--impl reference: the CPU restatement of the reference's commit (no Rust toolchain in this image, so the
    crate itself cannot run; kind = "port"), all host threads, bounded row sample per step."""
    if rank != 0:
        return
    nv = args.nv
    row_len, num_rows, cw = shape_for(nv)
    threads = os.cpu_count() or 1
    evals = gen_evals(nv, 0)
    # size the sample so that a step is ~1-2 s of CPU work
    rate, _, _, _ = cpu_commit_sample(nv, max(threads, 8), threads, True, evals)
    sample_rows = int(min(num_rows, max(threads, 1.5 * rate / row_len)))
    sample_rows = max(threads, sample_rows // threads * threads)
    for _ in range(args.warmup):
        cpu_commit_sample(nv, sample_rows, threads, True, evals)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_commit_sample(nv, sample_rows, threads, True, evals)
    dt = time.perf_counter() - t0
    value = args.steps * sample_rows * row_len / dt
    tuned, _, _, _ = cpu_commit_sample(nv, sample_rows, threads, False, evals)
    sample = f"rows 0..{sample_rows} of {num_rows} ({sample_rows * row_len} evals) per step, all rows+layers+roots produced"
    line = {
        "impl": "reference", "metric": "zip_commit_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (Int<4>)",
        "data": "synthetic",
        "config": {"workload": workload_name(nv), "nv": nv, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample,
                         "variant": "faithful (both permutations regenerated per row, as code_raa.rs:98-102)",
                         "tuned_value": tuned},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_full_commit_extrapolated": 1e3 * (1 << nv) / value,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--nv", type=int, default=24, help="log2 of the MLE size committed per GPU per step")
    ap.add_argument("--impl", default="zinc_b200", choices=["zinc_b200", "reference"])
    ap.add_argument("--no-sweep", action="store_true", help="skip the per-size table (nv=16,20,24,26) at N=1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--kernels-only", action="store_true",
                    help="only the device-resident timed steps (for ncu captures): no e2e, sweep, microbench, cpu")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import ctypes as C

    import torch
    import torch.distributed as dist

    from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices
    from zinc_b200 import _native as nat

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the zinc_b200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on STDOUT when NCCL_DEBUG=VERSION/INFO is set in the environment; stdout must
        # carry exactly one JSON line, so keep NCCL quiet here (ZINC_KEEP_NCCL_DEBUG=1 leaves the setting alone)
        if not os.environ.get("ZINC_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    L = nat.lib()
    ctx = Context(local_rank)
    nv = args.nv
    row_len, num_rows, cw = shape_for(nv)
    depth = cw.bit_length() - 1
    n_evals = 1 << nv

    def make_code(cw_, row_len_):
        code = RaaCode.with_permutations(ZipTypes(), row_len_, 2, shuffle_seeded_indices(cw_, KECCAK_SEEDS[0]),
                                         shuffle_seeded_indices(cw_, KECCAK_SEEDS[1]))
        return code, code.native(ctx, 1, 4)

    code, hcode = make_code(cw, row_len)

    # ---- device-resident inputs and outputs (the `value` leg) ----
    evals_h = gen_evals(nv, rank)
    pinned = torch.empty(n_evals, dtype=torch.int64).pin_memory()
    pinned.numpy().view(np.uint64)[:] = evals_h
    d_evals = pinned.to(dev, non_blocking=False)
    d_rows = torch.empty(num_rows * cw * 4, dtype=torch.int64, device=dev)
    d_layers = torch.empty(num_rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device=dev)
    d_roots = torch.empty(num_rows * 32, dtype=torch.uint8, device=dev)
    # an explicit non-default stream: the kernels, the CUDA events and torch's copies all live on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0

    def step_device():
        nat.check(L.zipgpu_commit_device(hcode, num_rows, d_evals.data_ptr(), d_rows.data_ptr(), d_layers.data_ptr(),
                                         d_roots.data_ptr(), sptr))

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step_device()
    barrier()
    nat.check(L.zipgpu_profile_enable(ctx.handle, 1))
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - launches0
    enc_ms, hash_ms, calls = C.c_double(), C.c_double(), C.c_uint64()
    nat.check(L.zipgpu_profile_read(ctx.handle, C.byref(enc_ms), C.byref(hash_ms), C.byref(calls), 1))
    nat.check(L.zipgpu_profile_enable(ctx.handle, 0))
    ms_per_step = ms_total / args.steps
    value = world * n_evals / (ms_per_step * 1e-3)
    roots_first = d_roots.cpu().numpy().copy()
    # in the commit steps the library ran the fused commit kernel (encode + Merkle levels 0..log2(E)) followed by the
    # batched upper-level passes: its profile slots hold those two
    fused_ms = enc_ms.value / max(calls.value, 1)
    upper_ms = hash_ms.value / max(calls.value, 1)

    # ---- the two kernels on their own, same inputs, timed live with CUDA events inside the library:
    #      the RAA encoder (HBM roofline) and the BLAKE3 tree passes (INT32 alu-pipe roofline) ----
    def timed_profile(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        nat.check(L.zipgpu_profile_read(ctx.handle, None, None, None, 1))
        nat.check(L.zipgpu_profile_enable(ctx.handle, 1))
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        e_, h_, c_ = C.c_double(), C.c_double(), C.c_uint64()
        nat.check(L.zipgpu_profile_read(ctx.handle, C.byref(e_), C.byref(h_), C.byref(c_), 1))
        nat.check(L.zipgpu_profile_enable(ctx.handle, 0))
        return e_.value / max(c_.value, 1), h_.value / max(c_.value, 1)

    enc_only_ms, _ = timed_profile(
        lambda: nat.check(L.zipgpu_encode_rows_device(hcode, num_rows, d_evals.data_ptr(), d_rows.data_ptr(), sptr)),
        args.steps)
    _, hash_only_ms = timed_profile(
        lambda: nat.check(L.zipgpu_merkle_rows_device(ctx.handle, num_rows, depth, 4, d_rows.data_ptr(),
                                                      d_layers.data_ptr(), d_roots.data_ptr(), sptr)),
        args.steps)
    assert np.array_equal(d_roots.cpu().numpy(), roots_first), "two-kernel path and fused path disagree on the roots"
    enc_ms.value, hash_ms.value, calls.value = enc_only_ms, hash_only_ms, 1
    if args.kernels_only:
        if rank == 0:
            emit({"metric": "zip_commit_evals_per_sec", "value": value, "unit": "evals/s",
                  "ms_per_step": ms_per_step, "fused_kernel_ms": fused_ms, "upper_passes_ms": upper_ms,
                  "encode_only_ms": enc_only_ms, "hash_only_ms": hash_only_ms,
                  "gpu_launches": int(launches), "note": "kernels-only run"})
        return

    # ---- e2e leg: host buffers through the C ABI (pinned H2D of evals + D2H of roots inside the timed region) ----
    roots_h = torch.empty(num_rows * 32, dtype=torch.uint8).pin_memory()

    def step_e2e():
        h = C.c_void_p()
        nat.check(L.zipgpu_commit_resident(hcode, num_rows, pinned.data_ptr(), roots_h.data_ptr(), C.byref(h)))
        L.zipgpu_data_free(h)

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    ctx.sync()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.stop()
    e2e_value = world * n_evals * args.steps / e2e_s
    assert np.array_equal(roots_h.numpy(), roots_first), "e2e roots differ from the device-resident run"

    # ---- full-materialisation variant, reported for transparency (rows + layers copied back to the host) ----
    e2e_full = None
    if world == 1 and nv <= 24:
        rows_h = torch.empty(num_rows * cw * 4, dtype=torch.int64).pin_memory()
        layers_h = torch.empty(num_rows * ((2 << depth) - 2) * 32, dtype=torch.uint8).pin_memory()
        nat.check(L.zipgpu_commit(hcode, num_rows, pinned.data_ptr(), rows_h.data_ptr(), layers_h.data_ptr(),
                                  roots_h.data_ptr()))
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            nat.check(L.zipgpu_commit(hcode, num_rows, pinned.data_ptr(), rows_h.data_ptr(), layers_h.data_ptr(),
                                      roots_h.data_ptr()))
        dt = (time.perf_counter() - t0) / reps
        e2e_full = {"value": n_evals / dt, "unit": "evals/s", "ms_per_step": dt * 1e3,
                    "d2h_bytes_per_step": int(rows_h.numel() * 8 + layers_h.numel() + num_rows * 32),
                    "note": "zipgpu_commit with rows+layers+roots all copied back to pinned host memory (PCIe-bound)"}
        del rows_h, layers_h

    # ---- INT32 peaks for the hasher's roofline ----
    alu, mix = C.c_double(), C.c_double()
    nat.check(L.zipgpu_microbench_int32(ctx.handle, 0, 2000, C.byref(alu)))
    nat.check(L.zipgpu_microbench_int32(ctx.handle, 1, 2000, C.byref(mix)))

    # ---- per-size table (device-resident, kernel-only), N = 1 ----
    sizes = {}
    if world == 1 and not args.no_sweep:
        for snv in (16, 20, 24, 26):
            srl, snr, scw = shape_for(snv)
            sdepth = scw.bit_length() - 1
            try:
                if snv == nv:
                    shc, sev, srows, slay, sroots = hcode, d_evals, d_rows, d_layers, d_roots
                else:
                    _, shc = make_code(scw, srl)
                    sev = torch.from_numpy(gen_evals(snv, 0).view(np.int64)).to(dev)
                    srows = torch.empty(snr * scw * 4, dtype=torch.int64, device=dev)
                    slay = torch.empty(snr * ((2 << sdepth) - 2) * 32, dtype=torch.uint8, device=dev)
                    sroots = torch.empty(snr * 32, dtype=torch.uint8, device=dev)
                run = lambda: nat.check(L.zipgpu_commit_device(shc, snr, sev.data_ptr(), srows.data_ptr(),
                                                               slay.data_ptr(), sroots.data_ptr(), sptr))
                enc = lambda: nat.check(L.zipgpu_encode_rows_device(shc, snr, sev.data_ptr(), srows.data_ptr(), sptr))
                reps = 10 if snv <= 24 else 4
                out = {}
                for name, fn in (("commit_ms", run), ("encode_ms", enc)):
                    for _ in range(3):
                        fn()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    a.record(stream)
                    for _ in range(reps):
                        fn()
                    b.record(stream)
                    torch.cuda.synchronize()
                    out[name] = a.elapsed_time(b) / reps
                out["evals_per_s"] = (1 << snv) / (out["commit_ms"] * 1e-3)
                out["encode_GBps"] = ENC_BYTES_PER_EVAL * (1 << snv) / (out["encode_ms"] * 1e-3) / 1e9
                sizes[f"nv{snv}"] = out
                if snv != nv:
                    del sev, srows, slay, sroots
                    torch.cuda.empty_cache()
            except Exception as ex:  # e.g. not enough memory for nv=26 next to the main buffers
                sizes[f"nv{snv}"] = {"error": str(ex)[:200]}

    # ---- the reference's own criterion shapes (benches/zip_benches.rs:225-262): EncodeRows / Commit at 2^12..2^16,
    #      EncodeMessage (one row, row_len 128..4096), MerkleRoot (one tree, 2^12..2^16 leaves); device-resident GPU
    #      time per call next to the CPU port on all host threads (these are latency-bound on a GPU: one row = one CTA) ----
    ref_shapes = None
    if rank == 0 and world == 1 and not args.no_sweep:
        try:
            def gpu_ms(fn, reps=50):
                for _ in range(5):
                    fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record(stream)
                for _ in range(reps):
                    fn()
                b.record(stream)
                torch.cuda.synchronize()
                return a.elapsed_time(b) / reps

            def cpu_ms(fn, reps=3):
                best = None
                for _ in range(reps):
                    t0 = time.perf_counter()
                    fn()
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                return best * 1e3

            cb = None
            if not args.no_cpu:
                from oracle import cbind as cb
                cb.build()
            threads = os.cpu_count() or 1
            ref_shapes = {"EncodeRows": {}, "Commit": {}, "EncodeMessage": {}, "MerkleRoot": {},
                          "note": "ms per call; gpu = device-resident buffers, CUDA events; cpu = oracle port, "
                                  f"{threads} threads (one thread for the single-row / single-tree cases)"}
            for P_ in (12, 13, 14, 15, 16):
                rl, nr, cw_ = shape_for(P_)
                dp = cw_.bit_length() - 1
                _, hc = make_code(cw_, rl)
                ev_h = gen_evals(P_, 0)
                ev = torch.from_numpy(ev_h.view(np.int64)).to(dev)
                rows_ = torch.empty(nr * cw_ * 4, dtype=torch.int64, device=dev)
                lay_ = torch.empty(nr * ((2 << dp) - 2) * 32, dtype=torch.uint8, device=dev)
                roots_ = torch.empty(nr * 32, dtype=torch.uint8, device=dev)
                g_enc = gpu_ms(lambda: nat.check(L.zipgpu_encode_rows_device(hc, nr, ev.data_ptr(), rows_.data_ptr(), sptr)))
                g_com = gpu_ms(lambda: nat.check(L.zipgpu_commit_device(hc, nr, ev.data_ptr(), rows_.data_ptr(),
                                                                        lay_.data_ptr(), roots_.data_ptr(), sptr)))
                ref_shapes["EncodeRows"][f"2^{P_}"] = {"gpu_ms": g_enc}
                ref_shapes["Commit"][f"2^{P_}"] = {"gpu_ms": g_com}
                # the call a host makes: pinned host evaluations in, host roots out, prover data resident
                ev_pin = torch.from_numpy(ev_h.view(np.int64)).pin_memory()
                roots_pin = torch.empty(nr * 32, dtype=torch.uint8).pin_memory()

                def host_call():
                    hd = C.c_void_p()
                    nat.check(L.zipgpu_commit_resident(hc, nr, ev_pin.data_ptr(), roots_pin.data_ptr(), C.byref(hd)))
                    L.zipgpu_data_free(hd)

                for _ in range(5):
                    host_call()
                t0 = time.perf_counter()
                for _ in range(50):
                    host_call()
                ref_shapes["Commit"][f"2^{P_}"]["e2e_host_ms"] = (time.perf_counter() - t0) / 50 * 1e3
                if cb is not None:
                    p1, p2 = cb.perm_from_seed(cw_, KECCAK_SEEDS[0]), cb.perm_from_seed(cw_, KECCAK_SEEDS[1])
                    ref_shapes["EncodeRows"][f"2^{P_}"]["cpu_ms_one_thread_tuned"] = cpu_ms(
                        lambda: cb.encode_rows(ev_h, nr, rl, 2, p1, p2))
                    ref_shapes["Commit"][f"2^{P_}"]["cpu_ms"] = cpu_ms(lambda: cb.commit_mt(
                        ev_h, nr, rl, 2, KECCAK_SEEDS[0], KECCAK_SEEDS[1], p1, p2, threads=threads, faithful=True))
                # MerkleRoot: ONE tree over 2^P random Int<4> leaves (zip_benches.rs:80-98)
                leaves_h = np.random.Generator(np.random.PCG64(P_)).integers(0, 1 << 64, size=(1 << P_) * 4, dtype=np.uint64)
                leaves = torch.from_numpy(leaves_h.view(np.int64)).to(dev)
                tl = torch.empty(((2 << P_) - 2) * 32, dtype=torch.uint8, device=dev)
                tr = torch.empty(32, dtype=torch.uint8, device=dev)
                ref_shapes["MerkleRoot"][f"2^{P_}"] = {"gpu_ms": gpu_ms(lambda: nat.check(L.zipgpu_merkle_rows_device(
                    ctx.handle, 1, P_, 4, leaves.data_ptr(), tl.data_ptr(), tr.data_ptr(), sptr)))}
                if cb is not None:
                    ref_shapes["MerkleRoot"][f"2^{P_}"]["cpu_ms"] = cpu_ms(lambda: cb.merkle_tree(P_, leaves_h, 4))
            for rl in (128, 256, 512, 1024, 2048, 4096):  # EncodeMessage: encode_wide of one row (zip_benches.rs:61-78)
                cw_ = 2 * rl
                _, hc = make_code(cw_, rl)
                msg_h = gen_evals(12, 0)[:rl].copy()
                msg = torch.from_numpy(msg_h.view(np.int64)).to(dev)
                out_ = torch.empty(cw_ * 4, dtype=torch.int64, device=dev)
                ent = {"gpu_ms": gpu_ms(lambda: nat.check(L.zipgpu_encode_rows_device(hc, 1, msg.data_ptr(), out_.data_ptr(), sptr)))}
                if cb is not None:
                    p1, p2 = cb.perm_from_seed(cw_, KECCAK_SEEDS[0]), cb.perm_from_seed(cw_, KECCAK_SEEDS[1])
                    ent["cpu_ms_faithful"] = cpu_ms(lambda: cb.encode_row_seeded(msg_h, 2, KECCAK_SEEDS[0], KECCAK_SEEDS[1]))
                    ent["cpu_ms_tuned"] = cpu_ms(lambda: cb.encode_rows(msg_h, 1, rl, 2, p1, p2))
                ref_shapes["EncodeMessage"][f"row_len={rl}"] = ent
        except Exception as ex:
            ref_shapes = {"error": str(ex)[:300]}

    # ---- the other LinearCode of the reference, ZipLinearCode (sparse code, zip/code.rs:77-215), same nv, same
    #      buffers: 0/1 matrices with row_len/2 cells per row (every coefficient 1, the densest the reference samples),
    #      tensor-core roofline for its GEMM and the CPU port on a bounded sample (rank 0, N = 1) ----
    sparse_code = None
    if rank == 0 and world == 1 and not args.no_sweep:
        try:
            from zinc_b200 import SparseMatrixZ, ZipLinearCode, ZipTypes
            srng = np.random.default_rng(0x5A00 + nv)

            def sample_matrix():
                d_ = row_len // 2
                cols = np.empty((cw // 2, d_), dtype=np.uint32)
                for i_ in range(cw // 2):
                    cols[i_] = np.sort(srng.permutation(row_len)[:d_])
                return SparseMatrixZ(cw // 2, row_len, d_, cols, np.ones(cw // 2 * d_, dtype=np.int64))

            scode = ZipLinearCode.with_matrices(ZipTypes(), row_len, cw, sample_matrix(), sample_matrix())
            sh = scode.native(ctx, 1, 4)
            s_run = lambda: nat.check(L.zipgpu_commit_device(sh, num_rows, d_evals.data_ptr(), d_rows.data_ptr(),
                                                             d_layers.data_ptr(), d_roots.data_ptr(), sptr))
            s_enc = lambda: nat.check(L.zipgpu_encode_rows_device(sh, num_rows, d_evals.data_ptr(), d_rows.data_ptr(), sptr))
            res = {}
            for name, fn in (("commit_ms", s_run), ("encode_ms", s_enc)):
                for _ in range(3):
                    fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record(stream)
                for _ in range(10):
                    fn()
                b.record(stream)
                torch.cuda.synchronize()
                res[name] = a.elapsed_time(b) / 10
            s_run()
            torch.cuda.synchronize()
            sparse_roots = d_roots.cpu().numpy().copy()
            sparse_rows_head = d_rows[: 256 * cw * 4].cpu().numpy().view(np.uint64).copy()
            ops = 2.0 * cw * row_len * num_rows * 8  # u8 x u8 multiply-adds x 2, 8 byte planes per evaluation
            tops = ops / (res["encode_ms"] * 1e-3) / 1e12
            sparse_code = {
                "workload": f"ZipLinearCode commit, 2^{nv} Int<1> evals, two {cw // 2}x{row_len} 0/1 matrices, "
                            f"{row_len // 2} cells per row", "kernel": scode.kernel_kind(ctx),
                "commit_ms": res["commit_ms"], "encode_ms": res["encode_ms"],
                "evals_per_s": n_evals / (res["commit_ms"] * 1e-3),
                "roofline": {"kernel": "split_planes_kernel + sparse_umma_kernel", "bound": "tensor", "achieved": tops,
                             "unit": "TOP/s", "ops_per_launch": ops},
            }
            if not args.no_cpu:
                from oracle import cbind
                threads = os.cpu_count() or 1
                nr = min(num_rows, 256)
                ma, mb = scode.a, scode.b
                t0 = time.perf_counter()
                rc, orows, _, oroots = cbind.sparse_commit(evals_h[: nr * row_len], nr, row_len, ma.n, ma.d, ma.cols, ma.coef,
                                                           mb.cols, mb.coef, threads=threads)
                dt = time.perf_counter() - t0
                sparse_code["cpu_baseline"] = {
                    "value": nr * row_len / dt, "unit": "evals/s", "cores": threads, "kind": "port",
                    "sample": f"rows 0..{nr} of {num_rows} ({dt:.1f} s)",
                    "rows_and_roots_match_gpu": bool(rc == 0 and np.array_equal(orows, sparse_rows_head[: nr * cw * 4]) and
                                                     np.array_equal(oroots, sparse_roots[: nr * 32]))}
        except Exception as ex:
            sparse_code = {"error": str(ex)[:300]}

    # ---- CPU baseline on a bounded sample (rank 0, N = 1) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        rate, _, _, _ = cpu_commit_sample(nv, max(threads, 8), threads, True, evals_h)
        sample_rows = int(min(num_rows, max(threads, 12.0 * rate / row_len)))  # ~12 s of CPU work
        sample_rows = max(threads, sample_rows // threads * threads)
        cpu_rate, cpu_s, sample_rows, cpu_roots = cpu_commit_sample(nv, sample_rows, threads, True, evals_h)
        tuned_rate, _, _, _ = cpu_commit_sample(nv, min(sample_rows, 64 * threads), threads, False, evals_h)
        parity = bool(np.array_equal(cpu_roots, roots_first[: sample_rows * 32]))
        cpu_baseline = {
            "value": cpu_rate, "unit": "evals/s", "cores": threads, "kind": "port",
            "sample": f"rows 0..{sample_rows} of {num_rows} of the same MLE ({sample_rows * row_len} evals, "
                      f"{cpu_s:.1f} s), rows+layers+roots produced",
            "variant": "faithful: permutations regenerated per row like code_raa.rs:98-102",
            "tuned_value": tuned_rate,
            "ms_per_full_commit_extrapolated": 1e3 * n_evals / cpu_rate,
            "roots_match_gpu": parity,
        }

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        if sparse_code and "roofline" in sparse_code:
            # no measured int8 number exists: 2 x the measured dense bf16 rate (tcgen05 kind::i8 : kind::f16 = 2 : 1)
            i8_peak = 2.0 * float(peaks.get("bf16_tflops", 1631.7))
            sp_traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    tr = json.load(f).get("sparse_umma_kernel", {})
                if tr.get("nv") == nv:
                    sp_traffic = tr.get("dram_bytes")
            except Exception:
                pass
            sparse_code["roofline"].update({"peak": i8_peak, "frac": sparse_code["roofline"]["achieved"] / i8_peak,
                                            "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (burst)",
                                            "traffic": sp_traffic})
        enc_ms_avg = enc_ms.value / max(calls.value, 1)
        hash_ms_avg = hash_ms.value / max(calls.value, 1)
        enc_gbs = ENC_BYTES_PER_EVAL * n_evals / (enc_ms_avg * 1e-3) / 1e9
        compressions = num_rows * (2 * cw - 1)
        hash_rate = compressions / (hash_ms_avg * 1e-3)
        clk = sampler.summary()
        sm_hz = 1e6 * float(clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0)
        num_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        alu_peak = ALU_LANES_PER_CLK_PER_SM * num_sms * sm_hz  # lane-instructions/s the alu pipe can issue
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture (profiles/)
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                tr = json.load(f).get("raa_encode_kernel", {})
            if tr.get("nv") == nv:
                traffic = tr.get("dram_bytes")
        except Exception:
            pass
        line = {
            "metric": "zip_commit_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {
                "workload": workload_name(nv), "nv": nv, "polys_per_gpu_per_step": 1,
                "outputs": "rows (u-hat) + all Merkle layers + roots materialised in HBM every step",
                "l2": f"per-step working set {(n_evals * 200) >> 20} MiB >> 126 MB L2 (inputs_larger_than_l2)",
                "parallelism": f"rows/polys sharded over {world} GPU(s), no data-path collective",
            },
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": n_evals * 8,
                    "d2h_bytes_per_step": num_rows * 32, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": "zipgpu_commit_resident (host evals in, host roots out, rows+layers stay in HBM)"},
            "e2e_full_d2h": e2e_full,
            "gpu_launches": int(launches),
            "roofline": {
                "kernel": "raa_encode_kernel", "bound": "hbm", "achieved": enc_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": enc_gbs / hbm_peak, "traffic": traffic, "ms_per_launch": enc_ms_avg,
                "measured": "encode-only launches (zipgpu_encode_rows_device) of the same workload, CUDA events inside "
                            "the library on the launch stream, same process, right after the timed commit steps",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": ENC_BYTES_PER_EVAL * n_evals,
            },
            "commit_kernels": {
                "fused_commit_kernel_ms": fused_ms, "upper_merkle_passes_ms": upper_ms,
                "note": "what the timed commit steps launch: raa_encode_kernel<FUSE> (encode + BLAKE3 leaves + tree "
                        "levels 1..log2(E) from shared memory) and the batched passes for the levels above",
                "two_kernel_path_ms": enc_ms_avg + hash_ms_avg,
                # the fused kernel is INT32-alu-bound: 31/32 of the compressions at 480 alu lane-instructions each, plus
                # the encoder's own alu work (45 M warp-instructions at nv = 24, ncu source counters)
                "fused_alu_pipe_util": ((compressions * 31.0 / 32.0) * HASH_ALU_SASS_PER_COMPRESSION +
                                        (45e6 * 32 if nv == 24 else 0.0)) / (fused_ms * 1e-3) / alu_peak,
                "step_vs_alu_floor": ((compressions * HASH_ALU_SASS_PER_COMPRESSION + (45e6 * 32 if nv == 24 else 0.0))
                                      / alu_peak) / (ms_per_step * 1e-3),
            },
            "hasher": {
                "kernels": "merkle_subtree_kernel x passes (tree-only launches of the same rows)", "bound": "int32_alu",
                "ms_per_step": hash_ms_avg,
                "compressions_per_step": compressions, "compressions_per_s": hash_rate,
                "lane_instr_per_s_min": hash_rate * HASH_INSTR_PER_COMPRESSION,
                "microbench_3alu_1fma_lane_ops_per_s": alu.value, "microbench_4alu_3fma_lane_ops_per_s": mix.value,
                "alu_pipe_util_sass": hash_rate * HASH_ALU_SASS_PER_COMPRESSION / alu_peak,
                "alu_pipe_frac_of_floor": hash_rate * HASH_ALU_MIN_PER_COMPRESSION / alu_peak,
                "alu_pipe_peak_lane_instr_per_s": alu_peak,
                "alu_model": "alu pipe = 64 lanes/clk/SM x SMs x sampled SM clock; 480 alu lane-instr per compression "
                             "issued (SASS), 456 irreducible (xor + rotate + feed-forward)",
                "share_of_step": hash_ms_avg / (hash_ms_avg + enc_ms_avg) if hash_ms_avg + enc_ms_avg > 0 else None,
            },
            "cpu_baseline": cpu_baseline,
            "sparse_code": sparse_code,
            "reference_bench_shapes": ref_shapes,
            "sizes": sizes,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
