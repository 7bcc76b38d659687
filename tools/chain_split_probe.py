"""Probe: does splitting the role-graph path of a small shard into two half-batch chains on two streams (so that each
chain's partly filled last wave of tiles is filled by the other chain) shorten forward + backward?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import situation_recognition_b200 as S
from situation_recognition_b200 import parallel
from situation_recognition_b200.synthetic import make_batch, make_train_json

def main(B=768, D=2048, iters=30):
    enc = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    torch.manual_seed(0)
    m = S.FCGGNN(enc, D, backbone=None).cuda().train()
    flat = parallel.attach(m, flat_params=True)
    fv, fn, gv, gn = [x.cuda() for x in make_batch(enc, B, D, seed=1)]
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    h = B // 2

    def one():
        flat.zero()
        out = m.predict_nouns(fn, gv, B)
        m.nouns_loss(out, gn).backward()

    def two():
        flat.zero()
        cur = torch.cuda.current_stream()
        sa.wait_stream(cur); sb.wait_stream(cur)
        with torch.cuda.stream(sa):
            o0 = m.predict_nouns(fn[:h], gv[:h], h)
        with torch.cuda.stream(sb):
            o1 = m.predict_nouns(fn[h:], gv[h:], B - h)
        cur.wait_stream(sa); cur.wait_stream(sb)
        (m.nouns_loss(o0, gn[:h]) + m.nouns_loss(o1, gn[h:])).backward()

    for name, fn_ in (("one chain", one), ("two chains", two)):
        for _ in range(3):
            fn_()
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn_()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for eng in m._engines.values():
            eng._packed_key = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn_()
        for _ in range(5):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print("B=%d %s: noun path fwd+bwd (graph replay, incl. weight pack) %.3f ms" % (B, name, e0.elapsed_time(e1) / iters))

if __name__ == "__main__":
    for B in (768, 1536):
        main(B)
