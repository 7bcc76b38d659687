"""Dataset + sharded batch sampler, API-compatible with the reference's `utils/imsitu_loader.py`."""
import os

import torch
import torch.utils.data as data

from .parallel import shard_range


class imsitu_loader(data.Dataset):
    """imsitu_loader.py:5-23: item = (img_name, img[3,224,224], verb:int, labels[3,R])."""

    def __init__(self, img_dir, train_json, encoder, transform=None):
        self.img_dir = img_dir
        self.train_json = train_json
        self.imgs_names = list(train_json.keys())
        self.encoder = encoder
        self.transform = transform
        self.skip_images = False     # extension: annotations only (the caller holds cached backbone features)

    def __getitem__(self, index):
        img_name = self.imgs_names[index]
        annotations = self.train_json[img_name]
        verb, labels = self.encoder.encode(annotations)
        if self.skip_images:
            return img_name, torch.empty(0), verb, labels
        from PIL import Image
        img = Image.open(os.path.join(self.img_dir, img_name)).convert('RGB')
        img = self.transform(img)
        return img_name, img, verb, labels

    def __len__(self):
        return len(self.train_json)


class ShardedBatchSampler(data.Sampler):
    """Yields, for every GLOBAL batch of `global_batch` samples, this rank's contiguous slice of it -- the same split
    `nn.DataParallel` makes along dim 0 (sr.py:467-470), but with one process per GPU.  Every rank yields the same
    number of batches, so the per-step collectives stay aligned.  No sample is dropped; samples are only repeated when
    the last global batch holds fewer samples than there are ranks (every rank needs at least one).  The shards of the
    last batch may be unequal: `global_sizes()` tells the loss how many images the whole batch holds."""

    def __init__(self, n, global_batch, rank=0, world=1, shuffle=False, seed=0):
        if global_batch < world:
            raise ValueError("batch_size %d is smaller than the number of ranks %d" % (global_batch, world))
        self.n, self.gb, self.rank, self.world, self.shuffle, self.seed = n, global_batch, rank, world, shuffle, seed
        self.epoch = 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return (self.n + self.gb - 1) // self.gb

    def global_sizes(self):
        """Images of each GLOBAL batch, in iteration order (the denominator of the verb loss, model.py:184-185)."""
        return [max(min(self.gb, self.n - i0), self.world) for i0 in range(0, self.n, self.gb)]

    def __iter__(self):
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(self.n, generator=g).tolist()
        else:
            order = list(range(self.n))
        for i0 in range(0, self.n, self.gb):
            chunk = order[i0:i0 + self.gb]
            if len(chunk) < self.world:         # a tail smaller than the world: pad by repetition (all ranks stay in step)
                chunk = (chunk * self.world)[:self.world]
            lo, hi = shard_range(len(chunk), self.rank, self.world)
            yield chunk[lo:hi]
