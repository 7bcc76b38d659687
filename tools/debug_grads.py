"""Debug helper: per-tensor gradient errors of the CUDA path against the golden fixture / oracle (GPU box)."""
import json, os, sys
os.environ.setdefault("SRG_POISON_WS", "1")
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import situation_recognition_b200 as S
from oracle import ggnn_oracle as O
from situation_recognition_b200.synthetic import make_batch, make_train_json

def relmax(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()

def run(B, D, compact, R4=False):
    if R4:
        G = os.path.join(ROOT, "tests", "golden")
        ann = json.loads(str(np.load(os.path.join(G, "encoder_overfitting.npz"))["annotations_json"]))
        enc = S.imsitu_encoder(ann, verbose=False)
    else:
        enc = S.imsitu_encoder(make_train_json(seed=0), verbose=False)
    params = O.init_params(enc.get_num_verbs(), enc.get_num_roles(), enc.get_num_labels(), D, seed=0)
    fv, fn, gv, gn = make_batch(enc, B, D, seed=7)
    t, c = O.build_tables(enc.roles_per_verb, enc.verb_list, enc.role_list)
    m = S.FCGGNN(enc, D, backbone=None, precision="bf16")
    m.load_state_dict(params, strict=False)
    m = m.cuda().eval()
    m._engine_for(torch.device("cuda", 0)).set_compact_rows(compact)
    pv, pn, gpn = m(fv.cuda(), gv.cuda(), img_nouns=fn.cuda())
    (m.verb_loss(pv, gv.cuda()) + m.nouns_loss(pn, gn.cuda())).backward()
    (_, _, _), grads, (opv, opn, ogpn) = O.train_step_grads(params, fv, fn, gv, gn, t, c, enc.get_num_labels(),
                                                            pred_verbs=pv.argmax(-1).cpu())
    print("B=%d D=%d compact=%d R4=%s logits %.2e %.2e %.2e" % (B, D, compact, R4, relmax(pv, opv), relmax(pn, opn), relmax(gpn, ogpn)))
    for k, p in m.named_parameters():
        print("   %-28s %.3e" % (k, relmax(p.grad, grads[k])))

if __name__ == "__main__":
    for compact in (0, 1):
        run(5, 256, compact, R4=True)
    for compact in (0, 1):
        run(16, 256, compact)
