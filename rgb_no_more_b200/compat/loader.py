"""`dataset_selector`-compatible loaders for the B200 data path (boundary B2).

The reference's `datasets.dataset_selector` (datasets.py:445-582) returns `torch.utils.data.DataLoader`s whose worker
processes decode one JPEG at a time and run the DCT transforms on the CPU, yielding `([Y, CbCr], labels)`.  `B200Loader` has
the same surface the reference's loop uses -- `len()`, iteration yielding `(inputs, labels)`, `.sampler.set_epoch(epoch)`
(train.py:143), `.batch_size`, `.dataset` -- but `inputs` is a `DCTBatch` living on the GPU: file bytes are read by host
threads, Huffman-decoded by `rgbnm_jpeg_decode_batch` into a pinned ring (`feeder.JpegFeeder`), copied on a copy stream and
pushed through the fused K0 kernel (`transforms.FusedDCT`), i.e. what comes out is already the operand of the patch-embedding
Linear.  Sharding follows the reference samplers (`ddp.shard_indices`: DistributedSampler for train incl. padding and
`set_epoch` reshuffle, DistributedEvalSampler -- strided, unpadded -- for eval; datasets.py:533-541, custom_sampler.py:88)."""
from __future__ import annotations

import csv
import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import torch

from .. import ddp
from .. import feeder as FD
from .. import transforms as TF


class NoChroma:
    """Second element of a DCTBatch: the fused kernel has already merged the chroma planes into the embed input.  Survives the
    reference's `cbcr = inputs[1].to(device)` (benchmark.py:329-332) and is recognised by the B200 model classes (`_rgbnm_absent`)."""
    _rgbnm_absent = True

    def to(self, *a, **kw):
        return self


NO_CHROMA = NoChroma()


class DCTBatch:
    """What the loader yields in place of the reference's `[Y, CbCr]`: the fused kernel's output on the device.
    Unpacks like the reference pair (`y, cbcr = inputs` / `inputs[0]`, `inputs[1]`): `y` is the embed input, `cbcr` a NoChroma marker
    -- the B200 model classes accept exactly that (`model(x, NO_CHROMA)`, like `model(x, None)`)."""

    def __init__(self, x: torch.Tensor):
        self.x = x

    def __iter__(self):
        yield self.x
        yield NO_CHROMA

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.x, NO_CHROMA)[i]

    def to(self, device, **kw):
        return DCTBatch(self.x.to(device, **kw))


class _Sampler:
    """The one sampler method the reference's loop calls (train.py:143)."""

    def __init__(self):
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)


def read_index(indexfile: str) -> List[tuple]:
    """`Filepath,Label` CSV (assets/indexbase_val.csv:1; pd.read_csv(...).to_numpy() in datasets.py:266)."""
    rows = []
    with open(indexfile, newline="") as f:
        rd = csv.reader(f)
        next(rd)
        for r in rd:
            if r:
                rows.append((r[0], int(r[1])))
    return rows


class B200Loader:
    def __init__(self, rows: Sequence[tuple], basepath: str, transform, batch_size: int, device, train: bool,
                 shuffle: bool, rank: int = 0, world_size: int = 1, io_threads: int = 8, decode_threads: int = 0, hb: int = 64,
                 wb: int = 64, prefetch: int = 2):
        # `transform`: a FusedDCT or a zero-argument factory of one (built at the first iteration: constructing it needs the GPU)
        self.rows, self.basepath, self._tf = list(rows), basepath, transform
        self.batch_size, self.train, self.shuffle = int(batch_size), train, shuffle
        self.rank, self.world = max(rank, 0), max(world_size, 1)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.sampler = _Sampler()
        self.dataset = self.rows                       # len(loader.dataset) is used for reporting only
        self.hb, self.wb, self.prefetch = hb, wb, max(1, prefetch)
        self.io_threads, self.decode_threads = io_threads, decode_threads
        self._feeder: Optional[FD.JpegFeeder] = None
        self._io: Optional[ThreadPoolExecutor] = None

    @property
    def tf(self) -> TF.FusedDCT:
        if not isinstance(self._tf, TF.FusedDCT):
            self._tf = self._tf()
        return self._tf

    # -- reference DataLoader surface -----------------------------------------------------------------------------
    def _indices(self) -> List[int]:
        n = len(self.rows)
        if self.train:      # DistributedSampler(shuffle=shuffle, drop_last=False), default seed 0 + set_epoch
            return ddp.shard_indices(n, self.rank, self.world, train=True, epoch=self.sampler.epoch, seed=0, shuffle=self.shuffle)
        return ddp.shard_indices(n, self.rank, self.world, train=False)

    def __len__(self) -> int:
        if not self.rows:
            return 0
        n = -(-len(self.rows) // self.world) if self.train else len(range(self.rank, len(self.rows), self.world))
        return -(-n // self.batch_size)

    def _read(self, i: int) -> bytes:
        p = self.rows[i][0]
        if self.basepath:
            p = os.path.join(self.basepath, p)
        with open(p, "rb") as f:
            return f.read()

    def __iter__(self):
        if not self.rows:
            return
        if self._feeder is None:
            self._feeder = FD.JpegFeeder(self.device, self.batch_size, self.hb, self.wb, slots=self.prefetch + 1,
                                         nthreads=self.decode_threads)
            self._io = ThreadPoolExecutor(max_workers=self.io_threads)
        fd = self._feeder
        idx = self._indices()
        batches = [idx[i:i + self.batch_size] for i in range(0, len(idx), self.batch_size)]
        queued: List[tuple] = []

        def submit(b):
            ids = b + [b[-1]] * (self.batch_size - len(b))          # the ring has fixed-size slots: pad the last batch
            # plan first, decode second: one vectorised plan draw per batch (global torch CPU generator, seeded SEED + rank by
            # dataset_selector); the decoder stops after the last block row a crop window needs
            plans = self.tf.sample_plans_packed(self.batch_size, self.hb, self.wb)
            last = plans["crop_i"].astype("int32") + plans["crop_size"].astype("int32") - 1
            fd.submit(list(self._io.map(self._read, ids)), last_rows=last)
            queued.append((b, plans))
        nxt = 0
        try:
            while nxt < len(batches) and len(queued) < self.prefetch:
                submit(batches[nxt])
                nxt += 1
            while queued:
                b, plans = queued.pop(0)
                y, c, q, flags, slot = fd.get()
                plans["clamp_in"] = flags                            # known only after the decode (dequantised range check)
                x = self.tf.run(y, c, q, plans, needs_stats=bool(plans["needs_stats"].any()))
                fd.release(slot)
                if nxt < len(batches):
                    submit(batches[nxt])
                    nxt += 1
                labels = torch.tensor([self.rows[i][1] for i in b], dtype=torch.int64)
                yield DCTBatch(x[:len(b)]), labels
        finally:
            while fd.pending:                                    # abandoned iteration: drain the ring
                fd.release(fd.get()[4])
            queued.clear()

    def close(self) -> None:
        if self._feeder is not None:
            self._feeder.close()
            self._io.shutdown(wait=True)
            self._feeder = None


def dataset_selector(dataset="imagenet_dct", type="train", indexpath="", basepath="", batch_size=128, num_workers=4, shuffle=False,
                     trainval_split=-1, return_indices=False, distributed=False, rank=-1, world_size=-1, seed=None, ops_list=None,
                     num_ops=2, ops_magnitude=10, dtype=torch.bfloat16, device=None, subblock=True):
    """Same arguments and return structure as the reference's `datasets.dataset_selector` (datasets.py:445-582) for the DCT
    datasets.  `num_workers` becomes the number of host decode threads of this rank (the reference gives each rank
    `num_cpus // world_size` DataLoader workers, pipeline_utils.py:125).  `subblock` (not a reference argument: there the sub-block
    conversion lives in the model's patch embedding, plainvit.py:173-216) selects the operand K0 writes for `--no_subblock`."""
    if dataset[0:12] != "imagenet_dct":
        raise NotImplementedError(f"rgbnm: dataset '{dataset}' is outside the B200 hot path (SURVEY.md 8f)")
    dev = device if device is not None else max(rank, 0)
    rows = read_index(indexpath)
    r, w = (rank, world_size) if distributed else (0, 1)

    def make(rows_, kind, shuf):
        # the reference's non-split branch does not forward ops_magnitude (datasets.py:568): kept, magnitude falls back to 10
        def tf(mag=ops_magnitude):
            extra = {} if (subblock or dataset != "imagenet_dct") else {"subblock": False}      # SwinV2 always decomposes blocks
            return TF.get_transform(dataset, "train" if kind == "train" else "val", ops_list=ops_list if kind == "train" else None,
                                    num_ops=num_ops, ops_magnitude=mag, dtype=dtype, device=dev, **extra)
        return B200Loader(rows_, basepath, tf, batch_size, dev, train=(kind == "train"), shuffle=shuf, rank=r, world_size=w,
                          decode_threads=max(1, num_workers))
    if trainval_split > 0:
        # identical split to the reference: torch.manual_seed(seed) then random_split over the index (datasets.py:514-521)
        g_state = torch.random.get_rng_state()
        if seed is not None:
            torch.manual_seed(seed)
        n = len(rows)
        vallen = int(trainval_split * n)
        trainlen = n - vallen
        tr, va = torch.utils.data.random_split(range(n), (trainlen, vallen))
        train_idx, val_idx = list(tr.indices), list(va.indices)
        if seed is not None:
            torch.manual_seed(seed + max(rank, 0))            # datasets.py:523: different augmentation streams per rank
        else:
            torch.random.set_rng_state(g_state)
        train_rows = [rows[i] for i in train_idx]
        out = (make(train_rows, "train", shuffle), make([rows[i] for i in val_idx], "val", False),
               make(train_rows[:round(trainlen * 0.05)], "val", False))
        return out if not return_indices else (*out, train_idx, val_idx)
    if type == "train":
        ops_magnitude = 10                                     # see `make`: datasets.py:568 drops the argument
    return make(rows, type if type == "train" else "val", shuffle)
