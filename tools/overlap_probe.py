"""Probe: do the independent chains of a step really overlap on the GPU?  Forward-only (no_grad, train mode) time of the
three paths with and without the side streams, eager vs CUDA graph, at a small shard size."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import situation_recognition_b200 as S
from situation_recognition_b200.synthetic import make_batch, make_train_json

def main(B=768, D=2048, iters=30):
    enc = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    torch.manual_seed(0)
    m = S.FCGGNN(enc, D, backbone=None).cuda().train()
    fv, fn, gv, gn = [x.cuda() for x in make_batch(enc, B, D, seed=1)]
    def fwd():
        with torch.no_grad():
            return m(fv, gv, img_nouns=fn)
    for overlap in (False, True):
        m.overlap_streams = overlap
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fwd()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            out = fwd()
        for _ in range(5):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print("B=%d overlap=%s forward (3 paths, graph replay): %.3f ms" % (B, overlap, e0.elapsed_time(e1) / iters))

if __name__ == "__main__":
    for B in (768, 1536):
        main(B)
