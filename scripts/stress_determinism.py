#!/usr/bin/env python
"""scripts/stress_determinism.py -- run-to-run bit-equality of the commit paths (the stand-in for a race checker:
compute-sanitizer is closed on this pool).  Repeats fused / two-kernel / host-pipelined commits of one polynomial and
compares rows, layers and roots by hash every time; any data race in the shared-memory phases, the warp-local write-out,
the bulk stores or the stream choreography shows up as a mismatch."""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch

    from helpers import KECCAK_SEEDS, shape_for
    from zinc_b200 import Context, RaaCode, ZipTypes, shuffle_seeded_indices
    from zinc_b200 import _native as nat

    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    L = nat.lib()
    ctx = Context(0)
    bad = 0
    for nv in (16, 20, 22, 24):
        row_len, num_rows, cw = shape_for(nv)
        depth = cw.bit_length() - 1
        code = RaaCode.with_permutations(ZipTypes(), row_len, 2, shuffle_seeded_indices(cw, KECCAK_SEEDS[0]),
                                         shuffle_seeded_indices(cw, KECCAK_SEEDS[1]))
        h = code.native(ctx, 1, 4)
        evals = np.random.default_rng(nv).integers(0, 1 << 64, size=1 << nv, dtype=np.uint64)
        pinned = torch.from_numpy(evals.view(np.int64)).pin_memory()
        d_ev = pinned.cuda()
        d_rows = torch.empty(num_rows * cw * 4, dtype=torch.int64, device="cuda")
        d_lay = torch.empty(num_rows * ((2 << depth) - 2) * 32, dtype=torch.uint8, device="cuda")
        d_roots = torch.empty(num_rows * 32, dtype=torch.uint8, device="cuda")
        roots_h = torch.empty(num_rows * 32, dtype=torch.uint8).pin_memory()

        def digest():
            torch.cuda.synchronize()
            # device-side checksums keep the run short: sums in int64 of every buffer + the roots themselves
            return (int(d_rows.sum().item()), int(d_lay.view(torch.int64).sum().item()),
                    hashlib.sha256(d_roots.cpu().numpy().tobytes()).hexdigest())

        ref = None
        for mode in ("fused", "two-kernel", "forced-fused"):
            os.environ.pop("ZIPGPU_NO_FUSE", None)
            os.environ.pop("ZIPGPU_FUSE_MIN_ROWS", None)
            if mode == "two-kernel":
                os.environ["ZIPGPU_NO_FUSE"] = "1"
            if mode == "forced-fused":
                os.environ["ZIPGPU_FUSE_MIN_ROWS"] = "1"
            for r in range(reps):
                d_rows.zero_(); d_lay.zero_(); d_roots.zero_()
                torch.cuda.synchronize()  # the library launches on its own non-blocking stream
                nat.check(L.zipgpu_commit_device(h, num_rows, d_ev.data_ptr(), d_rows.data_ptr(), d_lay.data_ptr(),
                                                 d_roots.data_ptr(), None))
                ctx.sync()
                got = digest()
                if ref is None:
                    ref = got
                elif got != ref:
                    bad += 1
                    print(f"MISMATCH nv={nv} mode={mode} rep={r}: {got} != {ref}")
        os.environ.pop("ZIPGPU_NO_FUSE", None)
        os.environ.pop("ZIPGPU_FUSE_MIN_ROWS", None)
        for r in range(reps):  # host-pipelined path: roots only
            hh = C.c_void_p()
            roots_h.zero_()
            nat.check(L.zipgpu_commit_resident(h, num_rows, pinned.data_ptr(), roots_h.data_ptr(), C.byref(hh)))
            L.zipgpu_data_free(hh)
            if hashlib.sha256(roots_h.numpy().tobytes()).hexdigest() != ref[2]:
                bad += 1
                print(f"MISMATCH nv={nv} host pipeline rep={r}")
        print(f"nv={nv}: {3 * reps} device commits + {reps} host commits identical" if not bad else f"nv={nv}: {bad} mismatches")
        del d_ev, d_rows, d_lay, d_roots
        torch.cuda.empty_cache()
    print("stress_determinism:", "OK" if not bad else f"{bad} MISMATCHES")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
