"""``SegmentData`` — the batch type the hot path consumes (mirror of reference ``data_utils/dataloader.py:27-53``).
The reference's own class is accepted too (duck-typed: ``.data`` dict of tensors, ``.segments`` list, ``.to``)."""
from __future__ import annotations

import dataclasses
import typing as tp

import torch


@dataclasses.dataclass
class SegmentData:
    data: tp.Dict[str, torch.Tensor]
    segments: tp.List[tp.Any]

    def __post_init__(self) -> None:
        if not isinstance(self.data, dict):
            raise TypeError(f"'features' need to be a dict, got: {type(self.data)}")
        if not self.data:
            raise ValueError(f"No data in {self}")
        if not isinstance(self.segments, list):
            raise TypeError(f"'segments' needs to be a list, got {self.segments}")
        batch_size = next(iter(self.data.values())).shape[0]
        if len(self.segments) != batch_size:
            raise RuntimeError(f"Incoherent batch size {batch_size} for {len(self.segments)} segments in {self}")

    def to(self, device: str) -> "SegmentData":
        out = {name: d.to(device) for name, d in self.data.items()}
        return SegmentData(data=out, segments=self.segments)

    def pin_memory(self) -> "SegmentData":
        return SegmentData(data={k: v.pin_memory() for k, v in self.data.items()}, segments=self.segments)

    def __getitem__(self, key: str) -> None:
        raise RuntimeError("New SegmentData batch is not a dict, use batch.data instead")


def synthetic_batch(batch_size=16, t=298, t_out=100, n_outputs=1000, n_subjects=4, seed=1234,
                    dims=(("text", 2, 3072), ("audio", 2, 1024), ("video", 2, 1408)), dtype=torch.float32, pin=False) -> SegmentData:
    """Synthetic window batch of the named shapes: N(0,1) features and fMRI, uniform subject ids.
    Same generator call order as ``oracle.tribe_oracle.synthetic_batch`` so both produce identical tensors."""
    g = torch.Generator().manual_seed(seed)
    data = {name: torch.randn(batch_size, l, d, t, generator=g).to(dtype) for name, l, d in dims}
    data["fmri"] = torch.randn(batch_size, n_outputs, t_out, generator=g)
    data["subject_id"] = torch.randint(0, n_subjects, (batch_size, 1), generator=g)
    batch = SegmentData(data=data, segments=[None] * batch_size)
    return batch.pin_memory() if pin else batch


class DevicePrefetcher:
    """Iterate over pinned host batches, issuing the host->device copy of batch i+1 on a side stream while the caller
    computes on batch i (the copy of a full TRIBE batch is 216 MB ~= 4 ms of PCIe time per step otherwise serialised
    in front of the forward).  Two persistent device slots are reused (no allocator traffic); a slot is overwritten
    only after the step that consumed it has finished (event recorded when the consumer asks for the next batch).
    Yields device-resident ``SegmentData``."""

    def __init__(self, batches, device=None):
        self.batches = batches
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [None, None]
        self.free = [None, None]

    def _load(self, batch, j):
        slot = self.slots[j]
        if slot is None or any(k not in slot or slot[k].shape != v.shape or slot[k].dtype != v.dtype for k, v in batch.data.items()):
            slot = self.slots[j] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in batch.data.items()}
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            if self.free[j] is not None:
                self.stream.wait_event(self.free[j])
            for k, v in batch.data.items():
                slot[k].copy_(v, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return SegmentData(data={k: slot[k] for k in batch.data}, segments=batch.segments), ready

    def feed(self, batches) -> "DevicePrefetcher":
        """Iterate over another stream of host batches through the SAME two device slots (stable device addresses are
        what lets ``graphed.GraphedTrainStep`` replay captured steps)."""
        self.batches = batches
        return self

    def resident(self, batches) -> list:
        """Load up to two host batches into the slots and return them as device-resident ``SegmentData``."""
        out = []
        for j, batch in enumerate(list(batches)[:2]):
            cur, ready = self._load(batch, j)
            torch.cuda.current_stream(self.device).wait_event(ready)
            out.append(cur)
        torch.cuda.synchronize(self.device)
        self.free = [None, None]
        return out

    def __iter__(self):
        it = iter(self.batches)
        try:
            nxt = self._load(next(it), 0)
        except StopIteration:
            return
        j = 0
        while nxt is not None:
            cur, ready = nxt
            main = torch.cuda.current_stream(self.device)
            main.wait_event(ready)
            try:
                nxt = self._load(next(it), j ^ 1)
            except StopIteration:
                nxt = None
            yield cur
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))  # the consumer's step on slot j is enqueued by now
            self.free[j] = done
            j ^= 1
