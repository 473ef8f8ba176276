"""Batch-mode front-end: the host side that feeds `FastSpeech2.forward` in `synthesize.py --mode batch`.

Reference: `dataset.TextDataset` (dataset.py:149-198) wrapped in `DataLoader(dataset, batch_size=8, collate_fn=dataset.collate_fn)`
(synthesize.py:191-198), each batch then moved by `utils.tools.to_device` (utils/tools.py:58-65).

* `TextDataset` keeps the reference's interface (constructor arguments, `__len__`, `__getitem__`, `collate_fn`, tuple layout, numpy dtypes),
  so the untouched `DataLoader(...)` line of the CLI works on it (`dropin.patch_text_dataset()` binds it as `dataset.TextDataset`).  The
  phoneme strings are converted once, at construction, and `collate_fn` pads into one preallocated array.
* `TextBatches` is the loader to use when the caller is free to choose: whole batches are prepared by a background thread (bounded
  queue), optionally length-bucketed (sorted by phoneme count before cutting into batches, which removes most of the padding the
  unmasked convolutions would otherwise compute on), sharded by whole batches across ranks like `parallel.shard_microbatches`, and
  `device_batches()` stages them through pinned memory on a copy stream so the host-to-device copy of batch i + 1 overlaps the
  forward of batch i.  It yields exactly what `to_device(collate_fn(...))` yields.

Text normalisation / G2P (`text/`, lexicons) is outside the hot path: the symbol-id function is taken
from the deployment's own `text` package, or passed in.
"""
from __future__ import annotations

import json
import os
import queue
import threading
from typing import Callable, Iterator, List, Optional, Sequence, Tuple

import numpy as np


def read_source(path: str) -> Tuple[List[str], List[str], List[str], List[str]]:
    """`name|speaker|{phonemes}|raw text` lines (the layout of `preprocessed_data/*/val.txt`, dataset.py:174-186).  A line without
    exactly four fields is an error, as in the reference (its tuple unpacking raises ValueError)."""
    names, speakers, texts, raws = [], [], [], []
    with open(path, "r", encoding="utf-8") as f:
        for no, line in enumerate(f, 1):
            fields = line.rstrip("\n").split("|")
            if len(fields) != 4:
                raise ValueError(f"{path}:{no}: expected 4 '|'-separated fields (name|speaker|text|raw_text), got {len(fields)}")
            names.append(fields[0]); speakers.append(fields[1]); texts.append(fields[2]); raws.append(fields[3])
    return names, speakers, texts, raws


def _default_text_to_sequence() -> Callable[[str, Sequence[str]], List[int]]:
    try:
        from text import text_to_sequence          # the deployment's reference tree (text/__init__.py:17)
    except Exception as e:                          # pragma: no cover - depends on the deployment
        raise ImportError("fastspeech2_b200.frontend needs the reference's `text` package on sys.path (run from the reference tree, "
                          "as synthesize.py does) or an explicit text_to_sequence=... callable") from e
    return text_to_sequence


def pad_batch(seqs: Sequence[np.ndarray]) -> np.ndarray:
    """Right-pad 1-D id arrays with 0 (the PAD symbol id) to the longest one: what `utils.tools.pad_1D` returns (utils/tools.py:265-275)."""
    if not seqs:
        raise ValueError("empty batch")
    out = np.zeros((len(seqs), max(len(s) for s in seqs)), dtype=np.result_type(*[s.dtype for s in seqs]))
    for i, s in enumerate(seqs):
        out[i, : len(s)] = s
    return out


class TextDataset:
    """Same contract as `dataset.TextDataset` (dataset.py:149-198); usable with `torch.utils.data.DataLoader` (map-style)."""

    def __init__(self, filepath, preprocess_config, text_to_sequence: Optional[Callable] = None):
        self.cleaners = preprocess_config["preprocessing"]["text"]["text_cleaners"]
        self.basename, self.speaker, self.text, self.raw_text = read_source(filepath)
        with open(os.path.join(preprocess_config["path"]["preprocessed_path"], "speakers.json")) as f:
            self.speaker_map = json.load(f)
        t2s = text_to_sequence or _default_text_to_sequence()
        self.phones = [np.array(t2s(t, self.cleaners)) for t in self.text]
        self.lengths = np.array([p.shape[0] for p in self.phones], dtype=np.int64)

    def __len__(self):
        return len(self.text)

    def __getitem__(self, idx):
        return (self.basename[idx], self.speaker_map[self.speaker[idx]], self.phones[idx], self.raw_text[idx])

    def process_meta(self, filename):
        return read_source(filename)

    @staticmethod
    def collate_fn(data):
        ids = [d[0] for d in data]
        speakers = np.array([d[1] for d in data])
        raw_texts = [d[3] for d in data]
        text_lens = np.array([d[2].shape[0] for d in data])
        texts = pad_batch([d[2] for d in data])
        return ids, raw_texts, speakers, texts, text_lens, max(text_lens)


def plan_batches(lengths: Sequence[int], batch_size: int, bucket: bool) -> List[List[int]]:
    """Utterance indices per batch.  `bucket=False`: file order, as `DataLoader(batch_size, shuffle=False)`.  `bucket=True`: utterances
    sorted by phoneme count (stable, longest first, so the largest workspace is allocated by the first batch) before cutting."""
    if batch_size <= 0:
        raise ValueError("batch_size must be positive")
    order = list(range(len(lengths)))
    if bucket:
        order.sort(key=lambda i: -int(lengths[i]))
    return [order[i: i + batch_size] for i in range(0, len(order), batch_size)]


def padded_fraction(lengths: Sequence[int], batches: Sequence[Sequence[int]]) -> float:
    """Share of the padded [B, Lmax] phoneme grid that is padding, over all batches."""
    grid = sum(len(b) * max(int(lengths[i]) for i in b) for b in batches)
    return 1.0 - float(sum(int(lengths[i]) for b in batches for i in b)) / max(grid, 1)


class TextBatches:
    """Iterable over collated batches `(ids, raw_texts, speakers, texts, text_lens, max_len)` of one source file.

    rank / world: whole batches are dealt to ranks in contiguous blocks (`parallel.shard_microbatches`): the batch composition, and with
    it every output, is the same for every GPU count."""

    def __init__(self, filepath, preprocess_config, batch_size: int = 8, bucket: bool = False, text_to_sequence: Optional[Callable] = None,
                 prefetch: int = 2, rank: int = 0, world: int = 1):
        from .parallel import shard_microbatches
        self.dataset = TextDataset(filepath, preprocess_config, text_to_sequence)
        self.all_batches = plan_batches(self.dataset.lengths, batch_size, bucket)
        self.batches = [self.all_batches[i] for i in shard_microbatches(len(self.all_batches), rank, world)]
        self.prefetch = max(1, int(prefetch))

    def __len__(self):
        return len(self.batches)

    def padded_fraction(self) -> float:
        return padded_fraction(self.dataset.lengths, self.all_batches)

    def collate(self, indices: Sequence[int]):
        return self.dataset.collate_fn([self.dataset[i] for i in indices])

    def __iter__(self) -> Iterator[tuple]:
        """Batches prepared by a background thread, at most `prefetch` ahead of the consumer."""
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def put(item) -> bool:                         # False: the consumer is gone
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            try:
                for b in self.batches:
                    if not put(("batch", self.collate(b))):
                        return
                put(("end", None))
            except BaseException as e:                  # surfaced in the consumer thread
                put(("error", e))

        t = threading.Thread(target=work, name="fs2-text-batches", daemon=True)
        t.start()
        try:
            while True:
                kind, payload = q.get()
                if kind == "end":
                    return
                if kind == "error":
                    raise payload
                yield payload
        finally:
            stop.set()

    def device_batches(self, device) -> Iterator[tuple]:
        """What `utils.tools.to_device(batch, device)` returns for every batch (utils/tools.py:58-65: int64 speakers / texts / lengths on the
        device, ids / raw texts / max length untouched).  On a CUDA device the copies go through per-slot pinned staging buffers on a
        side stream; the consumer's stream waits on the slot's event, so batch i + 1 is uploaded while batch i runs."""
        import torch
        device = torch.device(device)
        if device.type != "cuda":
            for ids, raw, spk, texts, lens, mx in self:
                yield ids, raw, torch.from_numpy(spk).long().to(device), torch.from_numpy(texts).long().to(device), torch.from_numpy(lens).to(device), mx
            return
        copy_stream = torch.cuda.Stream(device)
        n_slots = self.prefetch + 1
        pinned: List[dict] = [dict() for _ in range(n_slots)]
        consumed: List[Optional["torch.cuda.Event"]] = [None] * n_slots

        def stage(slot: int, name: str, arr: np.ndarray):
            src = torch.from_numpy(np.ascontiguousarray(arr)).long()
            buf = pinned[slot].get(name)
            if buf is None or buf.numel() < src.numel():
                buf = pinned[slot][name] = torch.empty(max(src.numel(), 1), dtype=torch.int64).pin_memory()
            view = buf[: src.numel()].view(src.shape)
            view.copy_(src)
            return view.to(device, non_blocking=True)

        def upload(slot: int, batch):
            ids, raw, spk, texts, lens, mx = batch
            if consumed[slot] is not None:
                consumed[slot].synchronize()                 # the pinned buffers of this slot were read by an earlier upload: it has finished
            with torch.cuda.stream(copy_stream):
                t = (stage(slot, "spk", spk), stage(slot, "texts", texts), stage(slot, "lens", lens))
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            consumed[slot] = ev
            return ids, raw, t, mx, ev

        it = iter(self)
        pending = []
        slot = 0
        for batch in it:
            pending.append(upload(slot, batch))
            slot = (slot + 1) % n_slots
            if len(pending) > self.prefetch:
                yield self._hand_over(pending.pop(0), device)
        while pending:
            yield self._hand_over(pending.pop(0), device)

    @staticmethod
    def _hand_over(item, device):
        import torch
        ids, raw, (spk, texts, lens), mx, ev = item
        cur = torch.cuda.current_stream(device)
        cur.wait_event(ev)
        for t in (spk, texts, lens):
            t.record_stream(cur)                             # allocated on the copy stream, used on the consumer's
        return ids, raw, spk, texts, lens, mx
