"""Restatement of the reference's AdaptiveVec encoders (sqz/src/vec.rs) -- TEST INFRASTRUCTURE: it produces the byte-exact raw
parts of each of the eight storage variants so the C ABI's decoder (csrc/adaptive.cu) can be checked against them.

  choose_storage        vec.rs:1086-1131  (size estimates :187-189, :327-333, :723-728, :830-835, :966-971)
  SimpleSparse          vec.rs:123-127, construct :191-214
  DenseW<u8|u16, u32>   vec.rs:660-664, construct :730-755
  Dense4                vec.rs:761-766, construct :837-888 (len / 2 + 1 bytes, low nibble first, 15 -> fallback)
  Dense3                vec.rs:895-900, construct :973-1022 (len / 21 + 1 u64 words, 3 bits per value, 7 -> fallback)
  CompressedIndexSparse vec.rs:222-227, construct :335-398 (dense_data over the values, index_bytes, block_starts)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

VARIANTS = ("D3", "D4", "D8", "D16", "V", "S3", "S4", "S8")  # enum order, vec.rs:1029-1053


@dataclass
class Parts:
    variant: str
    len: int
    dense: Optional[np.ndarray] = None            # u64 words (D3/S3), u8 (D4/S4/D8/S8), u16 (D16)
    fb_idx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    fb_val: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    index_bytes: Optional[np.ndarray] = None
    block_starts: Optional[np.ndarray] = None


def _dense(kind: str, length: int, values: np.ndarray, indexes: Optional[np.ndarray]):
    """Dense3 / Dense4 / DenseW::construct -> (data, fallback idx, fallback val)"""
    values = np.asarray(values, dtype=np.uint32)
    idx = np.arange(length, dtype=np.uint32) if indexes is None else np.asarray(indexes, dtype=np.uint32)
    thresh = {"3": 7, "4": 15, "8": 255, "16": 65535}[kind]
    over = values >= thresh
    code = np.where(over, thresh, values).astype(np.uint64)
    if kind == "3":
        data = np.zeros(length // 21 + 1, dtype=np.uint64)
        np.bitwise_or.at(data, idx // 21, code << ((idx % 21).astype(np.uint64) * np.uint64(3)))
    elif kind == "4":
        data = np.zeros(length // 2 + 1, dtype=np.uint8)
        np.bitwise_or.at(data, idx >> 1, (code << ((idx & 1).astype(np.uint64) * np.uint64(4))).astype(np.uint8))
    elif kind == "8":
        data = np.zeros(length, dtype=np.uint8)
        data[idx] = code.astype(np.uint8)
    else:
        data = np.zeros(length, dtype=np.uint16)
        data[idx] = code.astype(np.uint16)
    return data, idx[over].copy(), values[over].copy()


def _compressed_index(length: int, indexes: np.ndarray):
    """index_bytes and block_starts exactly as CompressedIndexSparse::construct builds them (vec.rs:347-390)"""
    total_blocks = (length + 255) // 256
    index_bytes, block_starts = [], []
    cur_block, cur_block_start = 0, 0
    for i in indexes.tolist():
        block, offset = i // 256, i % 256
        if block > cur_block:
            block_starts.append(cur_block_start)
            cur_block = block
            cur_block_start = len(index_bytes)
        while len(block_starts) < block:
            block_starts.append(len(index_bytes))
        index_bytes.append(offset)
    block_starts.append(cur_block_start)
    while len(block_starts) < total_blocks:
        block_starts.append(len(index_bytes))
    block_starts.append(len(index_bytes))
    return np.array(index_bytes, dtype=np.uint8), np.array(block_starts, dtype=np.uint32)


def estimate_sizes(length: int, values: np.ndarray) -> dict:
    values = np.asarray(values, dtype=np.uint32)
    nv = len(values)
    over = lambda t: int((values >= t).sum()) * 8
    d3 = lambda ln: (ln // 21 + 1) * 8 + over(7)
    d4 = lambda ln: ln // 2 + over(15)
    d8 = lambda ln: ln + over(255)
    d16 = lambda ln: ln * 2 + over(65535)
    s = lambda inner: inner(nv) + nv + (length // 256) * 4
    return {"D3": d3(length), "D4": d4(length), "D8": d8(length), "D16": d16(length), "S3": s(d3), "S4": s(d4), "S8": s(d8), "V": nv * 8}


def choose_storage(length: int, values: np.ndarray) -> str:  # vec.rs:1086-1131, including its quirk: the S8 and V branches
    e = estimate_sizes(length, values)                         # compare against min_size without lowering it (S8) / at all (V)
    opt, min_size = "D3", e["D3"]
    for name in ("D4", "D8", "D16", "S3", "S4"):
        if e[name] < min_size:
            opt, min_size = name, e[name]
    if e["S8"] < min_size:
        opt = "S8"
    if e["V"] < min_size:
        opt = "V"
    return opt


def encode(length: int, values, indexes, variant: Optional[str] = None) -> Parts:
    """AdaptiveVec::new (vec.rs:1135-1160); `variant` forces a storage class (tests cover all eight on the same data)."""
    values = np.asarray(values, dtype=np.uint32)
    indexes = np.asarray(indexes, dtype=np.uint32)
    variant = variant or choose_storage(length, values)
    if variant == "V":
        return Parts("V", length, None, indexes.copy(), values.copy())
    if variant in ("D3", "D4", "D8", "D16"):
        data, fi, fv = _dense(variant[1:], length, values, indexes)
        return Parts(variant, length, data, fi, fv)
    data, fi, fv = _dense(variant[1:], len(values), values, None)  # dense_data over the values (fallback indexed by entry number)
    ib, bs = _compressed_index(length, indexes)
    return Parts(variant, length, data, fi, fv, ib, bs)
