"""Input pipeline for the AGCN hot path: .npy memory map -> pinned staging -> asynchronous H2D with prefetch.

The reference feeds the model from ``MultiModalDataset`` / ``NumpyDatasetLoader`` (torch_src/dataset.py:15-58,
torch_src/loader.py:22-33) through a ``torch.utils.data.DataLoader`` with ``num_workers=0``
(torch_src/session/training.py:18-25): every sample is copied out of the memory map with ``np.array(data[index])``, collated
on the main thread, and moved with a synchronous pageable ``features.float().cuda()`` (torch_src/session/session.py:168-174)
while the GPU idles.  At > 1000 sequences/s that loop, not the model, is the bottleneck.

Here the same on-disk format -- ``<modality>_<split>_features.npy`` of shape (samples, M, T, V, C) next to
``<split>_labels.npy`` -- is read by a background thread that gathers a whole batch straight from the memory map into a pinned
staging buffer (one vectorised ``np.take`` per modality), starts the H2D copy on a side stream and hands the consumer device
tensors plus an event; up to ``depth`` batches are in flight.  The sample order is the reference's: the batches come from
torch's own ``RandomSampler`` / ``SequentialSampler`` + ``BatchSampler``, so under the same seed the same indices are drawn as
by ``DataLoader(dataset, batch_size, shuffle=..., drop_last=...)``.  Iteration yields ``(features, labels, indices)`` like the
reference's loader: ``features`` is a tensor for one modality and a dict for several (dataset.py:42-49), already fp32 on the
device (so the session's ``.float().cuda()`` is a no-op), ``labels`` int64 on the device.
"""
import os
import queue
import threading
from typing import Dict, Iterator, Optional, Sequence, Union

import numpy as np
import torch
from torch.utils.data import BatchSampler, RandomSampler, SequentialSampler


class FeatureStore:
    """The reference's ``MultiModalDataset`` (torch_src/dataset.py:15-58) without the per-sample Python loop: memory-mapped
    feature arrays keyed by modality (file name up to the first underscore) and the label vector of one split."""

    def __init__(self, input_paths: Union[str, Sequence[str]], split: str, in_memory: bool = False, debug: bool = False):
        paths = [input_paths] if isinstance(input_paths, str) else [p[0] if isinstance(p, (tuple, list)) else p for p in input_paths]
        if not paths:
            raise ValueError("Must at least specify one data path")
        self.labels = np.load(os.path.join(paths[0], f"{split}_labels.npy"))
        self.features: Dict[str, np.ndarray] = {}
        for path in paths:
            for entry in sorted(os.scandir(path), key=lambda e: e.name):
                if entry.is_file() and "features" in entry.name and split in entry.name:
                    self.features[entry.name[:entry.name.index("_")]] = np.load(entry.path, mmap_mode=None if in_memory else "r")
        if not self.features:
            raise FileNotFoundError(f"no <modality>_{split}_features.npy under {paths}")
        if debug:                                   # DebuggingSession: first 100 samples (dataset.py:33-34)
            self.labels = self.labels[:100]
        for name, arr in self.features.items():
            if len(arr) < len(self.labels):
                raise ValueError(f"{name}: {len(arr)} samples but {len(self.labels)} labels")

    def __len__(self):
        return len(self.labels)

    def get_input_shape(self) -> dict:
        return {k: tuple(v.shape[1:]) for k, v in self.features.items()}

    def get_num_classes(self) -> int:
        return len(np.unique(self.labels))

    @classmethod
    def from_arrays(cls, features: Dict[str, np.ndarray], labels: np.ndarray):
        """A store over arrays that are already in memory (synthetic data, tests): {modality: (samples, M, T, V, C)} and the labels."""
        self = cls.__new__(cls)
        self.labels = np.asarray(labels)
        self.features = {k: np.asarray(v) for k, v in features.items()}
        for name, arr in self.features.items():
            if len(arr) < len(self.labels):
                raise ValueError(f"{name}: {len(arr)} samples but {len(self.labels)} labels")
        return self

    @classmethod
    def from_dataset(cls, dataset):
        """Adopts the arrays of a reference ``MultiModalDataset`` instance (duck-typed: ``labels_data`` and ``features_data``
        = {modality: (loader, array)}), so the drop-in launcher can swap the loader without touching the dataset code."""
        self = cls.__new__(cls)
        self.labels = np.asarray(dataset.labels_data)
        self.features = {k: v[1] for k, v in dataset.features_data.items()}
        for name, arr in self.features.items():
            if not isinstance(arr, np.ndarray):
                raise TypeError(f"{name}: only numpy-backed modalities can be prefetched (got {type(arr).__name__})")
        return self


class _Slot:
    def __init__(self, store: FeatureStore, batch: int, device: torch.device, pin: bool):
        self.host = {k: torch.empty((batch,) + tuple(v.shape[1:]), dtype=torch.from_numpy(np.empty(0, v.dtype)).dtype, pin_memory=pin)
                     for k, v in store.features.items()}
        self.host_labels = torch.empty((batch,), dtype=torch.int64, pin_memory=pin)
        self.dev = {k: torch.empty_like(v, device=device) for k, v in self.host.items()} if device.type == "cuda" else None
        self.dev_labels = torch.empty((batch,), dtype=torch.int64, device=device) if device.type == "cuda" else None
        self.ready = torch.cuda.Event() if device.type == "cuda" else None       # H2D of this slot finished
        self.consumed = torch.cuda.Event() if device.type == "cuda" else None    # consumer's stream is done with the device buffers


class PrefetchLoader:
    """``for features, labels, indices in PrefetchLoader(store, batch_size, shuffle=True, drop_last=True, device="cuda")``.

    The tensors of a batch are views of a ring of ``depth`` device buffers: they stay valid until the NEXT batch is requested (work
    already enqueued on the current stream stays safe: the copy stream waits for it); clone them to keep them longer.
    ``len(loader)`` is the number of batches."""

    def __init__(self, store, batch_size: int, shuffle: bool = False, drop_last: bool = False, device="cuda", depth: int = 3,
                 generator: Optional[torch.Generator] = None, **unused_dataloader_kwargs):
        if not isinstance(store, FeatureStore):
            store = FeatureStore.from_dataset(store)
        self.store, self.batch_size, self.drop_last = store, int(batch_size), drop_last
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.depth = max(2, int(depth))
        self._generator = generator
        sampler = RandomSampler(range(len(store)), generator=generator) if shuffle else SequentialSampler(range(len(store)))
        self.batch_sampler = BatchSampler(sampler, self.batch_size, drop_last)
        self.dataset = store                      # DataLoader attribute the reference's progress code reads (len(loader.dataset))
        self._slots = None

    def __len__(self):
        return len(self.batch_sampler)

    def _gather(self, slot: _Slot, idx: np.ndarray):
        n = len(idx)
        order = np.argsort(idx, kind="stable")            # ascending file offsets for the memory map, scattered back into batch order
        for k, arr in self.store.features.items():
            dst = slot.host[k].numpy()
            if n > 1 and not np.all(order == np.arange(n)):
                dst[:n][order] = arr[idx[order]]
            else:
                np.take(arr, idx, axis=0, out=dst[:n])
        slot.host_labels[:n] = torch.from_numpy(self.store.labels[idx].astype(np.int64))

    def __iter__(self) -> Iterator:
        cuda = self.device.type == "cuda"
        if self._slots is None:
            self._slots = [_Slot(self.store, self.batch_size, self.device, pin=cuda) for _ in range(self.depth)]
        slots = self._slots
        # drawn here, on the caller's thread, with DataLoader's RNG consumption: its iterator first draws a base seed from the
        # (default) generator, then the RandomSampler draws the permutation seed -- so a seeded run visits the same indices
        torch.empty((), dtype=torch.int64).random_(generator=self._generator)
        batches = list(self.batch_sampler)
        done: "queue.Queue" = queue.Queue()
        free: "queue.Queue" = queue.Queue()
        for i in range(self.depth):
            free.put(i)
        copy_stream = torch.cuda.Stream(self.device) if cuda else None
        stop = threading.Event()

        def worker():
            try:
                if cuda:
                    torch.cuda.set_device(self.device)
                for b in batches:
                    i = free.get()
                    if stop.is_set():
                        return
                    slot = slots[i]
                    idx = np.asarray(b, dtype=np.int64)
                    self._gather(slot, idx)
                    if cuda:
                        with torch.cuda.stream(copy_stream):
                            copy_stream.wait_event(slot.consumed)          # the consumer finished reading this slot's device buffers
                            for k in slot.host:
                                slot.dev[k][:len(idx)].copy_(slot.host[k][:len(idx)], non_blocking=True)
                            slot.dev_labels[:len(idx)].copy_(slot.host_labels[:len(idx)], non_blocking=True)
                            slot.ready.record(copy_stream)
                    done.put((i, idx))
                done.put(None)
            except BaseException as exc:                   # noqa: BLE001 -- surfaced on the consumer thread
                done.put(exc)

        if cuda:
            for s in slots:
                s.consumed.record(torch.cuda.current_stream(self.device))
        thread = threading.Thread(target=worker, daemon=True)
        thread.start()
        prev = None
        try:
            while True:
                item = done.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                i, idx = item
                slot, n = slots[i], len(idx)
                if cuda:
                    cur = torch.cuda.current_stream(self.device)
                    cur.wait_event(slot.ready)
                    feats = {k: (v[:n] if v.dtype == torch.float32 else v[:n].float()) for k, v in slot.dev.items()}
                    labels = slot.dev_labels[:n]
                else:
                    feats = {k: v[:n].to(torch.float32, copy=True) for k, v in slot.host.items()}      # (CPU use: tests only)
                    labels = slot.host_labels[:n].clone()
                if prev is not None:                       # the batch handed out before this one may now be recycled
                    if cuda:
                        slots[prev].consumed.record(torch.cuda.current_stream(self.device))
                    free.put(prev)
                prev = i
                yield (next(iter(feats.values())) if len(feats) == 1 else feats), labels, torch.from_numpy(idx)
        finally:
            stop.set()
            free.put(-1)
            thread.join(timeout=5)
