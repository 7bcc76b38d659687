"""imSitu vocabulary / index tables, API-compatible with the reference's `utils/imsitu_encoder.py`.

Same constructor argument (the train-set dict), same attributes (`verb_list`, `role_list`, `label_list`,
`roles_per_verb`, `max_role_count`, `roles_to_verb_tensor_list`, `verb2role_encoding`, transforms) and the same
methods with the same results (reference file:line in each docstring).  The vocabulary is built with dicts
(O(N) instead of the reference's O(N * |list|) `list.index` scans) and the batch methods are vectorised; the GPU
path does not call them per element at all -- `device_tables()` hands the flat tables to libsrggnn
(`srg_set_tables`) and the gather/mask runs in a CUDA kernel.
"""
import numpy as np
import torch


class imsitu_encoder:
    def __init__(self, train_set, verbose=True):
        # imsitu_encoder.py:8-69 -- insertion-ordered vocabularies
        self.max_label_count = 3
        verb_ids, role_ids, label_ids = {}, {}, {}
        self.roles_per_verb = {}
        for annotations in train_set.values():
            verb = annotations["verb"]
            if verb not in verb_ids:
                verb_ids[verb] = len(verb_ids)
                self.roles_per_verb[verb] = []
            vroles = self.roles_per_verb[verb]
            for frame in annotations["frames"]:
                for role, label in frame.items():
                    if role not in role_ids:
                        role_ids[role] = len(role_ids)
                    if role not in vroles:
                        vroles.append(role)
                    if label not in label_ids:
                        label_ids[label] = len(label_ids)
        self.verb_list = list(verb_ids)
        self.role_list = list(role_ids)
        self.label_list = list(label_ids)
        self._verb_ids, self._role_ids, self._label_ids = verb_ids, role_ids, label_ids
        self.max_role_count = max((len(r) for r in self.roles_per_verb.values()), default=0)
        if verbose:
            print('train set stats: \n\t verb count:', len(self.verb_list),
                  '\n\t role count:', len(self.role_list),
                  '\n\t label count:', len(self.label_list),
                  '\n\t max role count:', self.max_role_count)

        V, R = len(self.verb_list), self.max_role_count
        table = np.full((V, R), len(self.role_list), dtype=np.int64)      # imsitu_encoder.py:71-89
        count = np.zeros(V, dtype=np.int64)
        for vi, verb in enumerate(self.verb_list):
            roles = self.roles_per_verb[verb]
            count[vi] = len(roles)
            table[vi, :len(roles)] = [role_ids[r] for r in roles]
        self._role_count = count
        self.roles_to_verb_tensor_list = torch.from_numpy(table)
        self.verb2role_encoding = self.get_verb2role_encoding()
        self._adj_table = None
        self._transforms = None

    # ---- torchvision transforms, built lazily (imsitu_encoder.py:17-36)
    def _make_transforms(self):
        if self._transforms is None:
            import torchvision as tv
            normalize = tv.transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
            train = tv.transforms.Compose([tv.transforms.Resize(224), tv.transforms.RandomCrop(224),
                                           tv.transforms.RandomHorizontalFlip(), tv.transforms.ToTensor(), normalize])
            dev = tv.transforms.Compose([tv.transforms.Resize(224), tv.transforms.CenterCrop(224),
                                         tv.transforms.ToTensor(), normalize])
            self._transforms = (normalize, train, dev)
        return self._transforms

    @property
    def normalize(self):
        return self._make_transforms()[0]

    @property
    def train_transform(self):
        return self._make_transforms()[1]

    @property
    def dev_transform(self):
        return self._make_transforms()[2]

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_transforms"] = None
        return state

    # ---- tables
    def get_verb2role_encoding(self):
        """imsitu_encoder.py:93-112: list of int64 [R] tensors, 1 for real roles, 0 for padding."""
        R = self.max_role_count
        return [(torch.arange(R) < int(n)).to(torch.int64) for n in self._role_count]

    def get_max_role_count(self):
        return self.max_role_count

    def get_num_verbs(self):
        return len(self.verb_list)

    def get_num_roles(self):
        return len(self.role_list)

    def get_num_labels(self):
        return len(self.label_list)

    def get_role_count(self, verb_id):
        """imsitu_encoder.py:158-159."""
        return int(self._role_count[int(verb_id)])

    def get_role_ids(self, verb_id):
        """imsitu_encoder.py:168-170."""
        return self.roles_to_verb_tensor_list[verb_id]

    def get_role_ids_batch(self, verbs):
        """imsitu_encoder.py:172-180 (CPU result, like the reference)."""
        idx = torch.as_tensor(verbs).detach().to("cpu", torch.int64).reshape(-1)
        return self.roles_to_verb_tensor_list[idx]

    def get_adj_matrix_noself(self, verb_ids):
        """imsitu_encoder.py:209-229 (CPU float32 [B, R, R], like the reference)."""
        if self._adj_table is None:
            R = self.max_role_count
            i = np.arange(R)[:, None]
            j = np.arange(R)[None, :]
            tab = np.zeros((R + 1, R, R), dtype=np.float32)
            for n in range(R + 1):
                tab[n] = ((i < n) & (j < n) & (i != j)) | ((i >= n) & (i == j))
            self._adj_table = torch.from_numpy(tab)
        idx = torch.as_tensor(verb_ids).detach().to("cpu", torch.int64).reshape(-1)
        counts = torch.from_numpy(self._role_count)[idx]
        return self._adj_table[counts]

    def get_verb2role_encoding_batch(self, verb_ids):
        """imsitu_encoder.py:231-240."""
        idx = torch.as_tensor(verb_ids).detach().to("cpu", torch.int64).reshape(-1)
        return torch.stack([self.verb2role_encoding[int(i)] for i in idx]).type(torch.FloatTensor)

    # ---- label encoding (imsitu_encoder.py:161-166,182-207)
    def encode(self, item):
        verb = self._verb_ids[item["verb"]]
        return verb, self.get_label_ids(item["verb"], item["frames"])

    def get_label_ids(self, verb, frames):
        roles = self.roles_per_verb[verb]
        unk = self._label_ids.get("UNK")
        out = np.full((len(frames), self.max_role_count), len(self.label_list), dtype=np.int64)
        for fi, frame in enumerate(frames):
            for ri, role in enumerate(roles):
                lid = self._label_ids.get(frame[role], unk)
                if lid is None:
                    raise ValueError("'UNK' is not in list")  # the reference raises here too (list.index)
                out[fi, ri] = lid
        return torch.from_numpy(out)

    # ---- flat tables for the CUDA gather/mask kernel
    def device_tables(self):
        """(verb2roles int32 [V*R], role_count int32 [V]) as contiguous numpy arrays for srg_set_tables."""
        return (np.ascontiguousarray(self.roles_to_verb_tensor_list.numpy().astype(np.int32).reshape(-1)),
                np.ascontiguousarray(self._role_count.astype(np.int32)))


def tables_from_encoder(encoder):
    """Flat int32 tables from ANY encoder exposing the reference API (e.g. the reference's own class)."""
    if hasattr(encoder, "device_tables"):
        return encoder.device_tables()
    V, R = encoder.get_num_verbs(), encoder.get_max_role_count()
    table = torch.as_tensor(encoder.roles_to_verb_tensor_list).numpy().astype(np.int32).reshape(V * R)
    count = np.array([encoder.get_role_count(v) for v in range(V)], dtype=np.int32)
    return np.ascontiguousarray(table), np.ascontiguousarray(count)
