"""Synthetic imSitu-shaped data (the real train.json / images / checkpoint are not available offline).

`make_train_json(seed)` reproduces the vocabulary SHAPE of imSitu (README.md:22-26 of the reference): 504 verbs,
190 roles, 2001 labels, at most 6 roles per verb, with a role-count histogram close to imSitu's (mean 3.54).
"""
import numpy as np
import torch

ROLE_HISTOGRAM = {1: 3, 2: 60, 3: 190, 4: 180, 5: 55, 6: 16}  # sums to 504


def make_train_json(seed=0, num_verbs=504, num_roles=190, num_labels=2001, images_per_verb=4):
    rng = np.random.RandomState(seed)
    counts = []
    for n, k in ROLE_HISTOGRAM.items():
        counts += [n] * k
    counts = np.array(counts[:num_verbs] if num_verbs <= len(counts) else counts + [3] * (num_verbs - len(counts)))
    rng.shuffle(counts)
    role_names = ["role%03d" % i for i in range(num_roles)]
    labels = ["", "UNK"] + ["n%08d" % i for i in range(num_labels - 2)]
    # every role name used at least once: deal them round-robin, then fill randomly
    verb_roles = []
    cursor = 0
    for n in counts:
        roles = []
        while len(roles) < n:
            cand = role_names[cursor % num_roles] if cursor < num_roles else role_names[rng.randint(num_roles)]
            cursor += 1
            if cand not in roles:
                roles.append(cand)
        verb_roles.append(roles)
    train = {}
    label_cursor = 0
    for vi, roles in enumerate(verb_roles):
        verb = "verb%03d" % vi
        for k in range(images_per_verb):
            frames = []
            for _ in range(3):
                frame = {}
                for r in roles:
                    if label_cursor < num_labels:       # guarantee every label appears once
                        lab = labels[label_cursor]
                        label_cursor += 1
                    else:
                        lab = labels[rng.randint(num_labels)]
                    frame[r] = lab
                frames.append(frame)
            train["%s_%d.jpg" % (verb, k)] = {"verb": verb, "frames": frames}
    return train


def make_batch(encoder, B, D=2048, seed=1234, device="cpu"):
    """Synthetic backbone features / verb ids / labels of the reference's shapes."""
    g = torch.Generator().manual_seed(seed)
    feat_v = (torch.randn(B, D, generator=g).abs() * 0.5)
    feat_n = (torch.randn(B, D, generator=g).abs() * 0.5)
    V, R, L = encoder.get_num_verbs(), encoder.get_max_role_count(), encoder.get_num_labels()
    gt_verb = torch.randint(0, V, (B,), generator=g)
    counts = torch.tensor([encoder.get_role_count(int(v)) for v in gt_verb])
    gt_nouns = torch.randint(0, L, (B, 3, R), generator=g)
    pad = torch.arange(R)[None, None, :] >= counts[:, None, None]
    gt_nouns = torch.where(pad.expand(B, 3, R), torch.full_like(gt_nouns, L), gt_nouns)
    return feat_v.to(device), feat_n.to(device), gt_verb.to(device), gt_nouns.to(device)
