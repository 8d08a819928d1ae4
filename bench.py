#!/usr/bin/env python
"""bench.py -- Zip commit throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--nv 24] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE Zip commit (RAA encode of every row + one BLAKE3 Merkle tree per row) of ONE synthetic
2^nv-evaluation Int<1> MLE.

  N = 1   the whole commit on one GPU.
  N > 1   BASELINE.json configs[2]: the SAME commit sharded by contiguous row range over the N GPUs ("strong" scaling:
          total work fixed).  Rank g commits rows [g*R/N, (g+1)*R/N) with zipgpu_commit_device_sharded: the kernel that
          produces its roots stores them into every peer's buffer over NVLink and its last CTA runs the publish/wait
          handshake, so the exchange (the only one on this path) is inside the timed region and every
          rank ends a step holding all R roots.  Inside the run the gathered roots are compared with a full
          single-GPU commit on rank 0 (`sharded_roots_equal_single_gpu`).  The old figure -- N independent commits, no
          exchange -- is kept under `weak`.

  value     evaluations/s of the whole job = 2^nv / (max-over-ranks device time per step), inputs resident in HBM
  e2e       the same through the C ABI with HOST buffers: every rank copies its slice of the evaluations from pinned
            host memory and all R roots come back to the host inside the timed region
            (zipgpu_commit_resident[_sharded]; rows/layers stay in HBM for `open`)
  roofline  the dominant kernel of the timed step, commit_ws_kernel (encode + BLAKE3 leaves + 4 tree levels, INT32-alu
            bound): irreducible alu-pipe lane-ops per launch / its CUDA-event time, against the alu-pipe peak measured in
            the same run with zipgpu_microbench_int32(kind 2)
  roofline_encoder   the encoder alone (HBM-bound): 72 algorithmic bytes per evaluation / its launch time
  cpu_baseline  the oracle's multithreaded restatement of the reference's CPU commit, all rows of the same MLE
  configs   driver-visible numbers for the other BASELINE configs and the "next" rows (nv = 20 e2e, the 64 x 2^18
            batch, column openings, combine_rows, the prover flow)

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
COMMIT_BYTES_PER_EVAL = 8.0 + 64.0 + 128.0  # + the Merkle layers: compulsory HBM bytes of a whole commit
# alu-pipe (LOP3/SHF/PRMT/IADD3: 64 lanes/clk/SM on B200, scratch/pipes.cu) lane-ops per BLAKE3 compression: the
# irreducible 7 rounds x 8 G x (4 xor + 4 rotate) + 8 feed-forward xors = 456 (adds go to the fma pipe as IMAD); what
# the SASS of the hash loop issues per compression is 480
HASH_ALU_MIN_PER_COMPRESSION = 456.0
HASH_ALU_SASS_PER_COMPRESSION = 480.0
ALU_LANES_PER_CLK_PER_SM = 64.0


# stdout must carry exactly ONE JSON line: everything any library prints on fd 1 (NCCL's version banner, build chatter)
# goes to stderr, and the result line is written to the saved descriptor
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


def shard_range(n: int, rank: int, world: int):
    base, rem = divmod(n, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def gen_evals(nv: int, salt: int = 0) -> np.ndarray:
    """uniform 64-bit words, like DenseMultilinearExtension::rand (poly_z/mle/dense.rs:140-145)"""
    rng = np.random.Generator(np.random.PCG64(0x21C0 + nv + 1000 * salt))
    return rng.integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)


def workload_name(nv: int) -> str:
    row_len, num_rows, cw = shape_for(nv)
    return (f"zip_commit nv={nv}: 2^{nv} Int<1> evals as {num_rows} rows x {row_len}, RAA rep=2 -> cw={cw} Int<4>, "
            f"{num_rows} BLAKE3 Merkle trees of depth {cw.bit_length() - 1}; perm seeds = fresh KeccakTranscript")


def config_for(nv: int, gpus: int) -> dict:
    """the `config` object, identical in both arms (the driver compares them)"""
    total = ((1 << nv) * 200) >> 20
    return {
        "workload": workload_name(nv), "nv": nv,
        "outputs": "rows (u-hat) + all Merkle layers + roots materialised every step",
        "l2": f"inputs_larger_than_l2: per-step working set {total} MiB in total, {total // max(gpus, 1)} MiB per GPU "
              f"at N={gpus} (> 126 MB L2); no explicit flush",
    }


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
def cpu_commit(nv: int, rows: int, threads: int, faithful: bool, evals: np.ndarray, reps: int = 1):
    """the oracle port on `threads` host threads over rows 0..rows of the MLE -> (evals/s, seconds, roots)"""
    from oracle import cbind

    cbind.build()
    row_len, num_rows, cw = shape_for(nv)
    rows = min(rows, num_rows)
    p1, p2 = cbind.perm_from_seed(cw, KECCAK_SEEDS[0]), cbind.perm_from_seed(cw, KECCAK_SEEDS[1])
    ev = np.ascontiguousarray(evals[: rows * row_len])
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        rc, _, _, roots = cbind.commit_mt(ev, rows, row_len, 2, KECCAK_SEEDS[0], KECCAK_SEEDS[1], p1, p2,
                                          threads=threads, faithful=faithful, want_rows=True, want_layers=True)
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    return rows * row_len / best, best, roots


def run_reference_arm(args, rank: int, world: int):
    """--impl reference: the CPU restatement of the reference's commit (no Rust toolchain in this image, so the crate
    itself cannot run; kind = "port"), all host threads, EVERY row of the same MLE in every step."""
    if rank != 0:
        return
    nv = args.nv
    row_len, num_rows, cw = shape_for(nv)
    threads = os.cpu_count() or 1
    evals = gen_evals(nv)
    for _ in range(args.warmup):
        cpu_commit(nv, num_rows, threads, True, evals)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_commit(nv, num_rows, threads, True, evals)
    dt = time.perf_counter() - t0
    value = args.steps * (1 << nv) / dt
    tuned, _, _ = cpu_commit(nv, num_rows, threads, False, evals)
    sample = f"all {num_rows} rows (2^{nv} evals) per step; rows + layers + roots produced"
    line = {
        "impl": "reference", "metric": "zip_commit_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": config_for(nv, args.gpus),
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample,
                         "variant": "faithful (both permutations regenerated per row, as code_raa.rs:98-102)",
                         "tuned_value": tuned},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
_HOST_BLOCKS = []  # zipgpu_host_alloc blocks stay alive for the life of the process


def host_tensor(n, dtype):
    """a torch CPU tensor of n elements over zipgpu_host_alloc memory: pinned AND backed by 2 MiB pages where the OS
    grants them.  (torch's pin_memory() blocks were, on some boxes and depending on what the process had allocated
    before, fed to the GPU at half the PCIe rate: nv = 20 host-to-host 0.55 instead of 0.26 ms.)"""
    import ctypes as C

    import torch

    from zinc_b200 import _native as nat

    item = torch.empty(0, dtype=dtype).element_size()
    p = C.c_void_p()
    nat.check(nat.lib().zipgpu_host_alloc(max(n * item, 1), C.byref(p)))
    _HOST_BLOCKS.append(p)
    buf = (C.c_uint8 * max(n * item, 1)).from_address(p.value)
    return torch.frombuffer(buf, dtype=dtype, count=n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--nv", type=int, default=24, help="log2 of the MLE size committed per step")
    ap.add_argument("--impl", default="zinc_b200", choices=["zinc_b200", "reference"])
    ap.add_argument("--no-sweep", action="store_true", help="skip the per-size / per-config tables at N=1")
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
    from zinc_b200.dist import PeerRoots

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the zinc_b200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
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
    lay_per_row = ((2 << depth) - 2) * 32
    num_sms = torch.cuda.get_device_properties(dev).multi_processor_count

    def make_code(cw_, row_len_):
        code = RaaCode.with_permutations(ZipTypes(), row_len_, 2, shuffle_seeded_indices(cw_, KECCAK_SEEDS[0]),
                                         shuffle_seeded_indices(cw_, KECCAK_SEEDS[1]))
        return code, code.native(ctx, 1, 4)

    code, hcode = make_code(cw, row_len)

    # ---- this rank's share of THE commit: rows [begin, begin + count) ----
    begin, count = shard_range(num_rows, rank, world)
    evals_h = gen_evals(nv)  # the same MLE on every rank
    my_evals = evals_h[begin * row_len:(begin + count) * row_len]
    pinned = host_tensor(max(my_evals.size, 1), torch.int64)
    pinned.numpy().view(np.uint64)[: my_evals.size] = my_evals
    d_evals = pinned.to(dev, non_blocking=False)
    d_rows = torch.empty(max(count, 1) * cw * 4, dtype=torch.int64, device=dev)
    d_layers = torch.empty(max(count, 1) * lay_per_row, dtype=torch.uint8, device=dev)
    d_roots = torch.empty(max(count, 1) * 32, dtype=torch.uint8, device=dev)
    # an explicit non-default stream: the kernels, the CUDA events and torch's copies all live on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0

    peer = PeerRoots(ctx, num_rows) if world > 1 else None
    gathered = [0]

    def step_device():
        if peer is None:
            nat.check(L.zipgpu_commit_device(hcode, num_rows, d_evals.data_ptr(), d_rows.data_ptr(), d_layers.data_ptr(),
                                             d_roots.data_ptr(), sptr))
        else:
            gathered[0] = peer.commit_device(hcode, begin, count, d_evals.data_ptr(), d_rows.data_ptr(),
                                             d_layers.data_ptr(), sptr)

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
    value = n_evals / (ms_per_step * 1e-3)
    # what the timed steps launched: the fused commit kernel (profile slot "encode") and the upper tree passes
    fused_ms = enc_ms.value / max(calls.value, 1)
    upper_ms = hash_ms.value / max(calls.value, 1)

    # ---- N > 1: the gathered roots against a full single-GPU commit on rank 0 (and on every rank against each other) ----
    roots_ok = None
    if peer is not None:
        peer.sync()
        mine_all = peer.tensor(gathered[0]).clone()
        ref = torch.empty(num_rows * 32, dtype=torch.uint8, device=dev)
        if rank == 0:
            d_all = torch.from_numpy(evals_h.view(np.int64)).to(dev)
            nat.check(L.zipgpu_commit_device(hcode, num_rows, d_all.data_ptr(), None, None, ref.data_ptr(), sptr))
            torch.cuda.synchronize()
            del d_all
        dist.broadcast(ref, 0)
        ok = torch.tensor([1 if torch.equal(ref, mine_all) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        roots_ok = bool(ok.item())
        assert roots_ok, "sharded commit: gathered roots differ from the single-GPU commit"
        roots_first = ref.cpu().numpy().copy()
    else:
        roots_first = d_roots.cpu().numpy().copy()

    if args.kernels_only:
        if rank == 0:
            emit({"metric": "zip_commit_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
                  "ms_per_step": ms_per_step, "fused_kernel_ms": fused_ms, "upper_passes_ms": upper_ms,
                  "gpu_launches": int(launches), "sharded_roots_equal_single_gpu": roots_ok, "note": "kernels-only run"})
        if world > 1:
            peer.close()
            dist.destroy_process_group()
        return

    # ---- e2e leg: host buffers through the C ABI (pinned H2D of this rank's evaluations + D2H of ALL roots inside the
    #      timed region) ----
    roots_h = host_tensor(num_rows * 32, torch.uint8)

    def step_e2e():
        h = C.c_void_p()
        if peer is None:
            nat.check(L.zipgpu_commit_resident(hcode, num_rows, pinned.data_ptr(), roots_h.data_ptr(), C.byref(h)))
        else:
            nat.check(L.zipgpu_commit_resident_sharded(hcode, peer.handle, begin, count, pinned.data_ptr(),
                                                       roots_h.data_ptr(), C.byref(h)))
        if h:
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
    e2e_value = n_evals * args.steps / e2e_s
    assert np.array_equal(roots_h.numpy(), roots_first), "e2e roots differ from the device-resident run"

    # ---- N > 1: the former weak-scaling figure (N independent full commits, no exchange), for continuity ----
    weak = None
    if world > 1:
        try:
            w_ev = torch.from_numpy(gen_evals(nv, rank + 1).view(np.int64)).to(dev)
            w_rows = torch.empty(num_rows * cw * 4, dtype=torch.int64, device=dev)
            w_lay = torch.empty(num_rows * lay_per_row, dtype=torch.uint8, device=dev)
            w_roots = torch.empty(num_rows * 32, dtype=torch.uint8, device=dev)
            run_w = lambda: nat.check(L.zipgpu_commit_device(hcode, num_rows, w_ev.data_ptr(), w_rows.data_ptr(),
                                                             w_lay.data_ptr(), w_roots.data_ptr(), sptr))
            for _ in range(3):
                run_w()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(10):
                run_w()
            b.record(stream)
            barrier()
            w_ms = max_over_ranks(a.elapsed_time(b)) / 10
            weak = {"value": world * n_evals / (w_ms * 1e-3), "unit": "evals/s", "ms_per_step": w_ms,
                    "note": f"{world} independent 2^{nv} commits, one per GPU, no exchange (round-1 headline)"}
            del w_ev, w_rows, w_lay, w_roots
        except Exception as ex:
            weak = {"error": str(ex)[:200]}

    extras = {}
    if world == 1:
        extras = single_gpu_extras(args, ctx, L, nat, C, torch, dev, stream, sptr, make_code, hcode, code, evals_h,
                                   pinned, d_evals, d_rows, d_layers, d_roots, roots_first, nv, num_sms)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        clk = sampler.summary()
        traffic_tab = {}
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of `ncu --set full` captures (profiles/)
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic_tab = json.load(f)
        except Exception:
            pass

        def traffic_of(name):
            tr = traffic_tab.get(name, {})
            return tr.get("dram_bytes") if tr.get("nv") == nv and world == 1 else None

        local_rows = count
        fused_levels = {8192: 4, 4096: 3, 2048: 3, 1024: 2, 512: 2}.get(cw, 0)
        comp_fused = local_rows * cw * (2.0 - 2.0 ** -fused_levels)  # leaves + levels 1..fused_levels
        comp_all = local_rows * (2 * cw - 1)
        line = {
            "metric": "zip_commit_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": config_for(nv, world),
            "parallelism": ("one GPU" if world == 1 else
                            f"ONE commit sharded by row range over {world} GPUs ({count} rows each); roots exchanged "
                            "inside the roots-producing kernel over NVLink peer memory (no NCCL on the data path)"),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(my_evals.size) * 8 * world,
                    "d2h_bytes_per_step": num_rows * 32 * world, "ms_per_step": 1e3 * e2e_s / args.steps,
                    "api": ("zipgpu_commit_resident" if world == 1 else "zipgpu_commit_resident_sharded") +
                           " (host evals in, host roots out, rows+layers stay in HBM)"},
            "gpu_launches": int(launches),
            "sharded_roots_equal_single_gpu": roots_ok,
            "weak": weak,
            "commit_kernels": {"fused_commit_kernel_ms": fused_ms, "upper_merkle_passes_ms": upper_ms,
                               "rows_per_gpu": local_rows},
        }
        alu_peak_model = ALU_LANES_PER_CLK_PER_SM * num_sms * 1e6 * float(clk.get("sm_mhz") or clk.get("sm_max_mhz") or 1965.0)
        alu_peak = extras.get("alu_pipe_peak_lane_ops_per_s") or alu_peak_model
        if fused_levels and fused_ms > 0:
            achieved = comp_fused * HASH_ALU_MIN_PER_COMPRESSION / (fused_ms * 1e-3)
            line["roofline"] = {
                "kernel": "commit_ws_kernel (RAA encode + BLAKE3 leaves + tree levels 1..%d, one launch)" % fused_levels,
                "bound": "int32_alu", "achieved": achieved / 1e12, "peak": alu_peak / 1e12, "unit": "Tlane-op/s",
                "frac": achieved / alu_peak, "ms_per_launch": fused_ms, "share_of_step": fused_ms / ms_per_step,
                "traffic": traffic_of("commit_ws_kernel"),
                "algorithmic_bytes_per_launch": (8.0 + 64.0 + 128.0 * (1.0 - 2.0 ** -(fused_levels + 1))) * count * row_len,
                "algorithmic_lane_ops_per_launch": comp_fused * HASH_ALU_MIN_PER_COMPRESSION,
                "model": "irreducible alu-pipe work: 456 lane-ops (xor + rotate + feed-forward) per BLAKE3 compression x "
                         "compressions of the launch; the encoder's own alu work is not counted as useful",
                "peak_source": ("zipgpu_microbench_int32 kind 2 (LOP3/SHF only), same run" if extras.get("alu_pipe_peak_lane_ops_per_s")
                                else "64 lanes/clk/SM x SMs x sampled SM clock (model)"),
                "peak_model_64_lanes": alu_peak_model / 1e12,
                "step_frac_of_alu_floor": (comp_all * HASH_ALU_MIN_PER_COMPRESSION / alu_peak) / (ms_per_step * 1e-3),
                "sass_alu_pipe_util": comp_fused * HASH_ALU_SASS_PER_COMPRESSION / (fused_ms * 1e-3) / alu_peak,
                "hbm": {"achieved_GBps": COMMIT_BYTES_PER_EVAL * count * row_len / (ms_per_step * 1e-3) / 1e9,
                        "peak_GBps": float(peaks.get("hbm_gbs", 6650.0)),
                        "note": "whole step as an HBM figure: not the binding roof"},
            }
        line.update({k: v for k, v in extras.items() if k != "alu_pipe_peak_lane_ops_per_s"})
        if "roofline_encoder" in line:
            r = line["roofline_encoder"]
            hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
            r.update({"peak": hbm_peak, "frac": r["achieved"] / hbm_peak, "traffic": traffic_of("raa_encode_kernel"),
                      "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"})
            if "roofline" not in line:
                line["roofline"] = dict(r)
        if line.get("sparse_code") and "roofline" in line["sparse_code"]:
            # no measured int8 number exists: 2 x the measured dense bf16 rate (tcgen05 kind::i8 : kind::f16 = 2 : 1)
            i8_peak = 2.0 * float(peaks.get("bf16_tflops", 1631.7))
            sr = line["sparse_code"]["roofline"]
            sr.update({"peak": i8_peak, "frac": sr["achieved"] / i8_peak, "traffic": traffic_of("sparse_umma_kernel"),
                       "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (burst)"})
        emit(line)
    if world > 1:
        dist.barrier()
        peer.close()
        dist.destroy_process_group()


def single_gpu_extras(args, ctx, L, nat, C, torch, dev, stream, sptr, make_code, hcode, code, evals_h, pinned, d_evals,
                      d_rows, d_layers, d_roots, roots_first, nv, num_sms):
    """N = 1 only: the kernels on their own (rooflines), the other BASELINE configs, the CPU baseline."""
    row_len, num_rows, cw = shape_for(nv)
    depth = cw.bit_length() - 1
    n_evals = 1 << nv
    out = {}

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

    def gpu_ms(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def wall_ms(fn, reps=10, warm=3, batches=3):
        """host wall time per call: best of `batches` batches of `reps` calls (the first batches after a change of
        buffer sizes still pay for allocations and page-locking; the best batch is the steady state)"""
        for _ in range(warm):
            fn()
        ctx.sync()
        best = None
        for _ in range(batches):
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            ctx.sync()
            dt = (time.perf_counter() - t0) / reps * 1e3
            best = dt if best is None else min(best, dt)
        return best

    # ---- the two kernels of the unfused path on their own, same inputs, timed with CUDA events inside the library ----
    enc_only_ms, _ = timed_profile(
        lambda: nat.check(L.zipgpu_encode_rows_device(hcode, num_rows, d_evals.data_ptr(), d_rows.data_ptr(), sptr)),
        args.steps)
    _, hash_only_ms = timed_profile(
        lambda: nat.check(L.zipgpu_merkle_rows_device(ctx.handle, num_rows, depth, 4, d_rows.data_ptr(),
                                                      d_layers.data_ptr(), d_roots.data_ptr(), sptr)),
        args.steps)
    assert np.array_equal(d_roots.cpu().numpy(), roots_first), "two-kernel path and fused path disagree on the roots"
    enc_gbs = ENC_BYTES_PER_EVAL * n_evals / (enc_only_ms * 1e-3) / 1e9
    out["roofline_encoder"] = {
        "kernel": "raa_encode_kernel", "bound": "hbm", "achieved": enc_gbs, "unit": "GB/s", "ms_per_launch": enc_only_ms,
        "measured": "encode-only launches (zipgpu_encode_rows_device) of the same workload, CUDA events inside the "
                    "library on the launch stream, same process, right after the timed commit steps",
        "algorithmic_bytes_per_launch": ENC_BYTES_PER_EVAL * n_evals,
    }
    if args.kernels_only:
        return out

    # ---- INT32 peaks: the alu pipe alone (the hash kernels' roofline) and the BLAKE3 mixes ----
    alu_only, alu3, mix = C.c_double(), C.c_double(), C.c_double()
    nat.check(L.zipgpu_microbench_int32(ctx.handle, 2, 2000, C.byref(alu_only)))
    nat.check(L.zipgpu_microbench_int32(ctx.handle, 0, 2000, C.byref(alu3)))
    nat.check(L.zipgpu_microbench_int32(ctx.handle, 1, 2000, C.byref(mix)))
    out["alu_pipe_peak_lane_ops_per_s"] = alu_only.value
    compressions = num_rows * (2 * cw - 1)
    out["hasher"] = {
        "kernels": "merkle_subtree_kernel x passes (tree-only launches of the same rows; the two-kernel path)",
        "bound": "int32_alu", "ms_per_step": hash_only_ms, "compressions_per_step": compressions,
        "compressions_per_s": compressions / (hash_only_ms * 1e-3),
        "frac_of_alu_peak": compressions * HASH_ALU_MIN_PER_COMPRESSION / (hash_only_ms * 1e-3) / alu_only.value,
        "microbench_alu_only_lane_ops_per_s": alu_only.value, "microbench_3alu_1fma_lane_ops_per_s": alu3.value,
        "microbench_8alu_6fma_lane_ops_per_s": mix.value, "two_kernel_path_ms": enc_only_ms + hash_only_ms,
    }

    if args.no_sweep:
        sizes = ref_shapes = sparse_code = None
        configs = {}
    else:
        configs = {}
        # ---- per-size table (device-resident, kernel-only) ----
        sizes = {}
        for snv in (16, 20, 22, 24, 26, 27):
            srl, snr, scw = shape_for(snv)
            sdepth = scw.bit_length() - 1
            try:
                if snv == nv:
                    shc, sev, srows, slay, sroots = hcode, d_evals, d_rows, d_layers, d_roots
                else:
                    _, shc = make_code(scw, srl)
                    sev = torch.from_numpy(gen_evals(snv).view(np.int64)).to(dev)
                    srows = torch.empty(snr * scw * 4, dtype=torch.int64, device=dev)
                    slay = torch.empty(snr * ((2 << sdepth) - 2) * 32, dtype=torch.uint8, device=dev)
                    sroots = torch.empty(snr * 32, dtype=torch.uint8, device=dev)
                run = lambda: nat.check(L.zipgpu_commit_device(shc, snr, sev.data_ptr(), srows.data_ptr(),
                                                               slay.data_ptr(), sroots.data_ptr(), sptr))
                enc = lambda: nat.check(L.zipgpu_encode_rows_device(shc, snr, sev.data_ptr(), srows.data_ptr(), sptr))
                reps = 10 if snv <= 24 else (4 if snv <= 26 else 2)
                o = {"commit_ms": gpu_ms(run, reps), "encode_ms": gpu_ms(enc, reps)}
                o["evals_per_s"] = (1 << snv) / (o["commit_ms"] * 1e-3)
                o["encode_GBps"] = ENC_BYTES_PER_EVAL * (1 << snv) / (o["encode_ms"] * 1e-3) / 1e9
                o["commit_vs_alu_floor"] = o["commit_ms"] * 1e-3 / (snr * (2 * scw - 1) * HASH_ALU_MIN_PER_COMPRESSION / alu_only.value)
                sizes[f"nv{snv}"] = o
                if snv == 20:
                    # BASELINE configs[1]: commit 2^20 on one B200, through the host API as well
                    ev_pin = host_tensor(1 << snv, torch.int64)
                    ev_pin.numpy()[:] = gen_evals(snv).view(np.int64)
                    roots_pin = host_tensor(snr * 32, torch.uint8)

                    def host_call():
                        hd = C.c_void_p()
                        nat.check(L.zipgpu_commit_resident(shc, snr, ev_pin.data_ptr(), roots_pin.data_ptr(), C.byref(hd)))
                        L.zipgpu_data_free(hd)

                    e2e20 = wall_ms(host_call, 20)
                    configs["commit_2^20_single_gpu"] = {
                        "device_ms": o["commit_ms"], "e2e_ms": e2e20, "evals_per_s_device": o["evals_per_s"],
                        "evals_per_s_e2e": (1 << snv) / (e2e20 * 1e-3), "h2d_bytes": 8 << snv, "d2h_bytes": snr * 32}
                if snv != nv:
                    del sev, srows, slay, sroots
                    torch.cuda.empty_cache()
            except Exception as ex:  # e.g. not enough memory for nv=26 next to the main buffers
                sizes[f"nv{snv}"] = {"error": str(ex)[:200]}

        # ---- BASELINE configs[3]: 64 independent 2^18 MLEs sharing one pp, zipgpu_batch_commit (host buffers) ----
        try:
            bnv, npoly = 18, 64
            brl, bnr, bcw = shape_for(bnv)
            _, bhc = make_code(bcw, brl)
            polys = [host_tensor(1 << bnv, torch.int64) for _ in range(npoly)]
            for i, pt in enumerate(polys):
                pt.numpy()[:] = gen_evals(bnv, 100 + i).view(np.int64)
            broots = [host_tensor(bnr * 32, torch.uint8) for _ in range(npoly)]
            ev_arr = (C.c_void_p * npoly)(*[p.data_ptr() for p in polys])
            rt_arr = (C.c_void_p * npoly)(*[r.data_ptr() for r in broots])
            bc = lambda: nat.check(L.zipgpu_batch_commit(bhc, npoly, bnr, ev_arr, None, None, rt_arr))
            b_ms = wall_ms(bc, 5, 2, 2)
            # the same batch device-resident, as one (64 * 512)-row matrix
            all_ev = torch.cat([p for p in polys]).to(dev)
            brows = torch.empty(npoly * bnr * bcw * 4, dtype=torch.int64, device=dev)
            blay = torch.empty(npoly * bnr * ((2 * bcw) - 2) * 32, dtype=torch.uint8, device=dev)
            brt = torch.empty(npoly * bnr * 32, dtype=torch.uint8, device=dev)
            bd_ms = gpu_ms(lambda: nat.check(L.zipgpu_commit_device(bhc, npoly * bnr, all_ev.data_ptr(), brows.data_ptr(),
                                                                    blay.data_ptr(), brt.data_ptr(), sptr)), 5)
            same = bool(np.array_equal(brt.cpu().numpy(), np.concatenate([r.numpy() for r in broots])))
            configs["batch_64x2^18"] = {
                "e2e_ms": b_ms, "device_ms": bd_ms, "evals_per_s_e2e": npoly * (1 << bnv) / (b_ms * 1e-3),
                "evals_per_s_device": npoly * (1 << bnv) / (bd_ms * 1e-3), "h2d_bytes": npoly * (8 << bnv),
                "d2h_bytes": npoly * bnr * 32, "api": "zipgpu_batch_commit (host pointers per polynomial, roots back)",
                "batch_roots_equal_one_matrix_commit": same}
            del all_ev, brows, blay, brt, polys, broots
            torch.cuda.empty_cache()
        except Exception as ex:
            configs["batch_64x2^18"] = {"error": str(ex)[:200]}

        # ---- f-1 / f-2: column openings (1000 columns, num_column_opening) and combine_rows on the resident data ----
        try:
            hd = C.c_void_p()
            roots_tmp = host_tensor(num_rows * 32, torch.uint8)
            nat.check(L.zipgpu_commit_resident(hcode, num_rows, pinned.data_ptr(), roots_tmp.data_ptr(), C.byref(hd)))
            ncols = 1000
            cols = np.random.default_rng(1).integers(0, cw, size=ncols, dtype=np.uint32)
            per = int(L.zipgpu_data_open_columns_wire_bytes(hd))
            wire = host_tensor(ncols * per, torch.uint8)
            ow = lambda: nat.check(L.zipgpu_data_open_columns_wire(hd, ncols, nat.ptr(cols), wire.data_ptr()))
            ow_ms = wall_ms(ow, 3, 1, 1)
            configs["open_columns_1000"] = {
                "ms": ow_ms, "bytes_out": ncols * per, "GBps": ncols * per / (ow_ms * 1e-3) / 1e9,
                "api": "zipgpu_data_open_columns_wire: 1000 columns x %d rows (entries + Merkle paths) as proof-stream "
                       "bytes, device gather + D2H to pinned host memory (PCIe-bound)" % num_rows}
            coeffs = np.random.default_rng(2).integers(0, 1 << 64, size=num_rows, dtype=np.uint64)
            comb = np.empty(row_len * 8, dtype=np.uint64)
            cr = lambda: nat.check(L.zipgpu_data_combine_rows(hd, nat.ptr(coeffs), 8, nat.ptr(comb)))
            cr_ms = wall_ms(cr, 10, 2)
            d_co = torch.from_numpy(coeffs.view(np.int64)).to(dev)
            d_out = torch.empty(row_len * 8, dtype=torch.int64, device=dev)
            crd_ms = gpu_ms(lambda: nat.check(L.zipgpu_combine_rows_device(ctx.handle, num_rows, row_len, d_evals.data_ptr(),
                                                                           d_co.data_ptr(), 8, d_out.data_ptr(), sptr)), 20)
            configs["combine_rows"] = {
                "device_ms": crd_ms, "host_call_ms": cr_ms, "GBps": 8.0 * n_evals / (crd_ms * 1e-3) / 1e9,
                "floor_ms_at_hbm_peak": 8.0 * n_evals / 6526.2e9 * 1e3,
                "note": "u' = sum_i coeff_i * row_i over Int<8>, 8 B read per evaluation (HBM-bound)"}
            L.zipgpu_data_free(hd)
        except Exception as ex:
            configs["open_and_combine"] = {"error": str(ex)[:200]}

        # ---- BASELINE configs[4]: the commit side of the prover flow (zinc/prover.rs:305-328) ----
        try:
            configs["prover_flow"] = prover_flow(ctx, args)
        except Exception as ex:
            configs["prover_flow"] = {"error": str(ex)[:300]}

        ref_shapes = reference_shapes(args, ctx, L, nat, C, torch, dev, stream, sptr, make_code, gpu_ms)
        sparse_code = sparse_section(args, ctx, L, nat, C, torch, dev, stream, sptr, evals_h, d_evals, d_rows, d_layers,
                                     d_roots, nv, gpu_ms)

    # ---- CPU baseline: the oracle port on all host threads, every row of the same MLE (rank 0, N = 1) ----
    cpu_baseline = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        cpu_rate, cpu_s, cpu_roots = cpu_commit(nv, num_rows, threads, True, evals_h)
        tuned_rate, _, _ = cpu_commit(nv, num_rows, threads, False, evals_h)
        cpu_baseline = {
            "value": cpu_rate, "unit": "evals/s", "cores": threads, "kind": "port",
            "sample": f"all {num_rows} rows of the same MLE (2^{nv} evals, {cpu_s:.1f} s), rows+layers+roots produced",
            "variant": "faithful: permutations regenerated per row like code_raa.rs:98-102",
            "tuned_value": tuned_rate, "ms_per_commit": 1e3 * cpu_s,
            "roots_match_gpu": bool(np.array_equal(cpu_roots, roots_first)),
        }
    out.update({"cpu_baseline": cpu_baseline, "configs": configs, "sizes": sizes, "sparse_code": sparse_code,
                "reference_bench_shapes": ref_shapes})
    return out


def prover_flow(ctx, args):
    """BASELINE configs[4] as far as this path goes: what ZincProver::commit_z_mle_and_prove_evaluation does around the
    commit (zinc/prover.rs:305-328) for z-vectors of 2^12..2^16 (spartan_benches.rs:37) and the 8-entry z of
    examples/simple_r1cs.rs -- RaaCode::new from a LIVE transcript (fresh seeds per proof, so the table build is on the
    critical path), setup, commit, then the open-side services on the resident data (1000 column openings as wire
    bytes, combine_rows).  Host-call wall times through the mirror; CPU port of the commit beside them."""
    from zinc_b200 import (DefaultLinearCodeSpec, DenseMultilinearExtension, KeccakTranscript, MultilinearZip, RaaCode)

    res = {}
    cb = None
    if not args.no_cpu:
        from oracle import cbind as cb

        cb.build()
    threads = os.cpu_count() or 1
    for nv in (3, 12, 14, 16):
        tr = KeccakTranscript()
        tr.absorb(b"prover_flow")  # a live transcript: seeds differ from the bench constants
        poly = DenseMultilinearExtension.rand(nv, np.random.default_rng(nv))
        reps = 5
        t = {"new_seeds_ms": 0.0, "new_perms_ms": 0.0, "new_tables_ms": 0.0, "setup_ms": 0.0, "commit_ms": 0.0,
             "open_1000_cols_ms": 0.0, "combine_rows_ms": 0.0}
        for rep in range(reps + 1):
            t0 = time.perf_counter()
            code = RaaCode.new(DefaultLinearCodeSpec(), 1 << nv, tr)  # two transcript.get_u64 (pure-Python Keccak here)
            ta = time.perf_counter()
            code.permutations()        # shuffle_seeded of 0..cw, twice (a Rust host: the rand crate)
            tb = time.perf_counter()
            code.native(ctx, 1, 4)     # zipgpu_code_create: edge-colouring, gather tables, upload -- once per proof
            t1 = time.perf_counter()
            pp = MultilinearZip.setup(1 << nv, code)
            t2 = time.perf_counter()
            data, comm = MultilinearZip.commit_resident(pp, poly, ctx)
            t3 = time.perf_counter()
            cw = code.codeword_len()
            cols = np.random.default_rng(rep).integers(0, cw, size=1000, dtype=np.uint32)
            wire = data.open_columns_wire_view(cols)  # proof-stream bytes in the context's pinned stream buffer
            t4 = time.perf_counter()
            co = np.random.default_rng(rep + 1).integers(0, 1 << 64, size=pp.num_rows, dtype=np.uint64)
            data.combine_rows(co, 8)
            t5 = time.perf_counter()
            data.free()
            ctx.drop_code(code)
            if rep:  # first repetition = warm-up
                for k, dt in zip(t, (ta - t0, tb - ta, t1 - tb, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
                    t[k] += dt * 1e3 / reps
        t["total_ms"] = sum(t.values())
        t["wire_bytes"] = len(wire)
        if cb is not None:
            row_len, num_rows = code.row_len(), pp.num_rows
            p1, p2 = code.permutations()
            ev = poly.evaluations.reshape(-1)
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                cb.commit_mt(ev, num_rows, row_len, 2, code.perm_1_seed, code.perm_2_seed, p1, p2,
                             threads=min(threads, num_rows), faithful=True)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            t["cpu_commit_ms"] = best * 1e3
            t["cpu_threads"] = min(threads, num_rows)
        res[f"z=2^{nv}"] = t
    res["note"] = ("host wall ms per call through the Python mirror (ctypes).  RaaCode::new is split: new_seeds = the two "
                   "transcript draws (pure-Python Keccak in this mirror; microseconds in Rust), new_perms = both shuffles, "
                   "new_tables = zipgpu_code_create (permutation check, edge-colouring, gather tables, upload)")
    return res


def reference_shapes(args, ctx, L, nat, C, torch, dev, stream, sptr, make_code, gpu_ms):
    """the reference's own criterion shapes (benches/zip_benches.rs:225-262): EncodeRows / Commit at 2^12..2^16,
    EncodeMessage (one row, row_len 128..4096), MerkleRoot (one tree, 2^12..2^16 leaves); device-resident GPU time per
    call next to the CPU port on all host threads (these are latency-bound on a GPU: one row = one CTA)"""
    try:
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
            ev_h = gen_evals(P_)
            ev = torch.from_numpy(ev_h.view(np.int64)).to(dev)
            rows_ = torch.empty(nr * cw_ * 4, dtype=torch.int64, device=dev)
            lay_ = torch.empty(nr * ((2 << dp) - 2) * 32, dtype=torch.uint8, device=dev)
            roots_ = torch.empty(nr * 32, dtype=torch.uint8, device=dev)
            g_enc = gpu_ms(lambda: nat.check(L.zipgpu_encode_rows_device(hc, nr, ev.data_ptr(), rows_.data_ptr(), sptr)), 50, 5)
            g_com = gpu_ms(lambda: nat.check(L.zipgpu_commit_device(hc, nr, ev.data_ptr(), rows_.data_ptr(),
                                                                    lay_.data_ptr(), roots_.data_ptr(), sptr)), 50, 5)
            ref_shapes["EncodeRows"][f"2^{P_}"] = {"gpu_ms": g_enc}
            ref_shapes["Commit"][f"2^{P_}"] = {"gpu_ms": g_com}
            # the call a host makes: pinned host evaluations in, host roots out, prover data resident
            ev_pin = host_tensor(ev_h.size, torch.int64)
            ev_pin.numpy()[:] = ev_h.view(np.int64)
            roots_pin = host_tensor(nr * 32, torch.uint8)

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
                ctx.handle, 1, P_, 4, leaves.data_ptr(), tl.data_ptr(), tr.data_ptr(), sptr)), 50, 5)}
            if cb is not None:
                ref_shapes["MerkleRoot"][f"2^{P_}"]["cpu_ms"] = cpu_ms(lambda: cb.merkle_tree(P_, leaves_h, 4))
        for rl in (128, 256, 512, 1024, 2048, 4096):  # EncodeMessage: encode_wide of one row (zip_benches.rs:61-78)
            cw_ = 2 * rl
            _, hc = make_code(cw_, rl)
            msg_h = gen_evals(12)[:rl].copy()
            msg = torch.from_numpy(msg_h.view(np.int64)).to(dev)
            out_ = torch.empty(cw_ * 4, dtype=torch.int64, device=dev)
            ent = {"gpu_ms": gpu_ms(lambda: nat.check(L.zipgpu_encode_rows_device(hc, 1, msg.data_ptr(), out_.data_ptr(), sptr)), 50, 5)}
            if cb is not None:
                p1, p2 = cb.perm_from_seed(cw_, KECCAK_SEEDS[0]), cb.perm_from_seed(cw_, KECCAK_SEEDS[1])
                ent["cpu_ms_faithful"] = cpu_ms(lambda: cb.encode_row_seeded(msg_h, 2, KECCAK_SEEDS[0], KECCAK_SEEDS[1]))
                ent["cpu_ms_tuned"] = cpu_ms(lambda: cb.encode_rows(msg_h, 1, rl, 2, p1, p2))
            ref_shapes["EncodeMessage"][f"row_len={rl}"] = ent
        return ref_shapes
    except Exception as ex:
        return {"error": str(ex)[:300]}


def sparse_section(args, ctx, L, nat, C, torch, dev, stream, sptr, evals_h, d_evals, d_rows, d_layers, d_roots, nv, gpu_ms):
    """the other LinearCode of the reference, ZipLinearCode (sparse code, zip/code.rs:77-215), same nv, same buffers:
    0/1 matrices with row_len/2 cells per row (every coefficient 1, the densest the reference samples), tensor-core
    roofline for its GEMM and the CPU port on a bounded sample"""
    row_len, num_rows, cw = shape_for(nv)
    n_evals = 1 << nv
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
        res = {"commit_ms": gpu_ms(s_run, 10), "encode_ms": gpu_ms(s_enc, 10)}
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
        return sparse_code
    except Exception as ex:
        return {"error": str(ex)[:300]}


if __name__ == "__main__":
    main()
