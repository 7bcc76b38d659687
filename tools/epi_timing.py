"""Where does the time of the fused GEMM epilogues go?  Debug build of libsrggnn.so with SM-clock counters in the
pipelined epilogue loop (csrc/gemm.cuh, -DSRG_EPI_TIMING); prints, per epilogue kind, the average clocks per tile and
per 32-column chunk spent waiting for the accumulator, for the chunk's TMA loads, for earlier TMA stores to release
their staging set, in tcgen05.ld + math, and issuing stores.

    SRG_NVCC_EXTRA=-DSRG_EPI_TIMING python tools/epi_timing.py [--batch 6144]
(the env var must also be set when building, so that the in-tree library is the instrumented one; rebuild without it
afterwards)."""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
assert "SRG_EPI_TIMING" in os.environ.get("SRG_NVCC_EXTRA", ""), "set SRG_NVCC_EXTRA=-DSRG_EPI_TIMING"

import situation_recognition_b200 as S  # noqa: E402
from situation_recognition_b200 import _lib, parallel  # noqa: E402
from situation_recognition_b200.synthetic import make_batch, make_train_json  # noqa: E402

KINDS = {2: "EPI_ZR", 3: "EPI_H", 5: "EPI_DRH", 6: "EPI_DH"}
COLS = ["tiles", "chunks", "tile_total", "wait_acc", "wait_in", "wait_stores+issue", "ld+math", "store_issue"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=6144)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    lib = _lib.load()
    fn = lib.srg_debug_epi_timing
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    enc = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    m = S.FCGGNN(enc, 2048, backbone=None, precision="bf16").cuda().train()
    flat = parallel.attach(m, flat_params=True)
    fv, fnn, gt_verb, gt_nouns = [x.cuda() for x in make_batch(enc, args.batch, 2048, seed=1)]

    def step():
        flat.zero()
        pv, pn, gpn = m(fv, gt_verb, img_nouns=fnn)
        (m.verb_loss(pv, gt_verb) + m.nouns_loss(pn, gt_nouns)).backward()

    buf = (ctypes.c_ulonglong * 64)()
    for _ in range(2):
        step()
    fn(buf, 1)
    for _ in range(args.steps):
        step()
    fn(buf, 1)
    mhz = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 0
    out = {"batch": args.batch, "sm_mhz_now": mhz, "kinds": {}}
    for k, name in KINDS.items():
        row = [int(buf[k * 8 + i]) for i in range(8)]
        tiles, chunks = max(row[0], 1), max(row[1], 1)
        out["kinds"][name] = {
            "warp_tiles": row[0], "chunks": row[1],
            "clk_per_tile_total": row[2] / tiles,
            "clk_per_tile_wait_acc": row[3] / tiles,
            "clk_per_chunk": {c: row[i] / chunks for i, c in enumerate(COLS) if i >= 4},
            "clk_per_chunk_busy": (row[2] - row[3]) / chunks,
        }
    print("EPI_TIMING " + json.dumps(out))


if __name__ == "__main__":
    main()
