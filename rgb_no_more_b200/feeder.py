"""JPEG-bytes -> device coefficient batches: the host stage in front of K0.

Replaces, for the DCT path, what the reference spreads over DataLoader worker processes, the default collate and
`utils.unpack_data`'s `.to(device)` (/root/reference/datasets.py:542-556, utils/pipeline_utils.py:52-76):

    JPEG bytes --rgbnm_jpeg_decode_batch (host threads, Huffman decode only)--> pinned staging slot
               --cudaMemcpyAsync on a copy stream--> device slot --K0--> (B,196,384) embed input

Decode of batch i+1 runs on the library's host threads (the ctypes call releases the GIL) while batch i's copy and the
GPU step are in flight; staging and device slots are reused round-robin, an event per slot orders reuse.  torch is used
for pinned memory, streams and events only."""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import torch

from . import dct_manip as dm
from . import lib as _lib


class JpegFeeder:
    def __init__(self, device, batch: int, hb: int = 64, wb: int = 64, slots: int = 3, nthreads: int = 0):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise _lib.RgbnmError("rgbnm: JpegFeeder needs a CUDA device; there is no CPU fallback")
        if slots < 2:
            raise ValueError("rgbnm: JpegFeeder needs at least two slots")
        self.batch, self.hb, self.wb, self.nthreads = batch, hb, wb, nthreads
        shapes = ((batch, hb, wb, 64), (batch, 2, hb // 2, wb // 2, 64), (batch, 3, 64))
        self.host = [tuple(torch.empty(s, dtype=torch.int16).pin_memory() for s in shapes) for _ in range(slots)]
        self.devb = [tuple(torch.empty(s, dtype=torch.int16, device=self.dev) for s in shapes) for _ in range(slots)]
        self.copied = [torch.cuda.Event() for _ in range(slots)]       # H2D of the slot finished (consumer waits on it)
        self.consumed = [torch.cuda.Event() for _ in range(slots)]     # consumer finished reading the device slot
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.pool = ThreadPoolExecutor(max_workers=1)
        self.pending: List = []            # (slot, future) in submission order
        self.next_slot = 0
        self.h2d_bytes = sum(t.numel() * 2 for t in self.host[0])

    def submit(self, jpegs: Sequence[bytes], last_rows=None) -> None:
        """Queue one batch of JPEG byte strings for decoding (returns immediately).  `last_rows`: plan-first decoding, see
        dct_manip.decode_batch (the plans of the batch are drawn before it is submitted)."""
        if len(jpegs) != self.batch:
            raise ValueError("rgbnm: JpegFeeder batches have a fixed size")
        slot = self.next_slot
        self.next_slot = (slot + 1) % len(self.host)
        if any(s == slot for s, _ in self.pending):
            raise RuntimeError("rgbnm: JpegFeeder ring is full; consume a batch before submitting another")
        # the previous H2D copy out of this pinned slot must be done before the decoder overwrites it
        self.copied[slot].synchronize()
        fut = self.pool.submit(dm.decode_batch, list(jpegs), self.hb, self.wb, self.nthreads, False, self.host[slot], last_rows)
        self.pending.append((slot, fut))

    def get(self, stream: Optional[torch.cuda.Stream] = None):
        """Oldest submitted batch on the device: (y, cbcr, quant, clamp_flags list, slot).  The returned tensors may be
        read on `stream` (default: current) and stay valid until `release(slot)` + `len(slots) - 1` further batches."""
        slot, fut = self.pending.pop(0)
        _, _, _, flags = fut.result()                   # decode errors surface here (RuntimeError, like dct_manip)
        stream = stream or torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])
            for dst, src in zip(self.devb[slot], self.host[slot]):
                dst.copy_(src, non_blocking=True)
            self.copied[slot].record(self.copy_stream)
        stream.wait_event(self.copied[slot])
        return (*self.devb[slot], flags.tolist(), slot)

    def release(self, slot: int, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Call after the last kernel that reads the slot's device tensors has been enqueued on `stream`."""
        self.consumed[slot].record(stream or torch.cuda.current_stream(self.dev))

    def close(self) -> None:
        self.pool.shutdown(wait=True)
