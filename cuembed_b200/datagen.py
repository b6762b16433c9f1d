"""Synthetic workload generator (host side, numpy).

Restates the reference's data recipe so that benchmark shapes mean the same
thing (file:line refer to NVIDIA/cuEmbed):
  * table values U(-1, 1) cast to the element type
    (utils/src/embedding_allocation.cu:113-116);
  * CSR bag lengths U{0..hotness} (:131-136);
  * lookup indices: per bag, draw until `hotness` DISTINCT categories from a
    power law over [1, N], N = num_categories - 1 (category 0 is never
    generated), y = floor((u * ((N+1)^g - 1) + 1)^(1/g)), g = 1 - alpha,
    u ~ U[0,1), optionally mapped through a random permutation of [0, N] and
    shuffled inside the bag (utils/src/datagen.cpp:39-50,86-132;
    utils/src/embedding_allocation.cu:139-158); alpha = 0 is uniform;
  * weights in {0.5, 0.25} (:160-168); grad_y integers in [-10, 10] (:234-237).
The reference's generators are std:: engines whose streams are
implementation-defined, so the exact arrays are not reproducible anyway; this
module uses numpy's PCG64 with fixed seeds and `double` arithmetic throughout
(the reference returns `float` from its inverse CDF, utils/src/datagen.cpp:40,
which quantises categories above 2^24).  The same arrays are given to the CPU
checker and to the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


def power_law_categories(rng: np.random.Generator, n: int, num_categories: int,
                         alpha: float) -> np.ndarray:
    """n draws from the power law over [1, num_categories] (int64)."""
    if alpha == 1.0:
        raise ValueError("alpha == 1 is singular (utils/src/datagen.cpp:44)")
    u = rng.random(n)
    if alpha == 0.0:
        y = u * num_categories + 1.0
    else:
        g = 1.0 - alpha
        hi = float(num_categories + 1) ** g
        y = (u * (hi - 1.0) + 1.0) ** (1.0 / g)
    y = np.floor(y).astype(np.int64)
    # u < 1 guarantees y <= num_categories up to rounding; clamp the edge.
    return np.clip(y, 1, num_categories)


def unique_bags(rng: np.random.Generator, batch_size: int, hotness: int,
                num_categories: int, alpha: float, permute: bool = True,
                shuffle: bool = True, perm: Optional[np.ndarray] = None
                ) -> np.ndarray:
    """[batch_size, hotness] int64 table rows, distinct inside every bag.

    Equivalent to the reference's "insert draws into a std::set until it holds
    `hotness` values" (utils/src/datagen.cpp:86-104): the kept values are the
    first `hotness` distinct values of an i.i.d. stream.
    """
    n_cat = num_categories - 1  # category 0 reserved
    if n_cat < hotness:
        raise ValueError("not enough categories for distinct bags")
    bags = power_law_categories(rng, batch_size * hotness, n_cat, alpha)
    bags = bags.reshape(batch_size, hotness)
    while True:
        bags.sort(axis=1)
        dup = np.zeros_like(bags, dtype=bool)
        dup[:, 1:] = bags[:, 1:] == bags[:, :-1]
        n_dup = int(dup.sum())
        if n_dup == 0:
            break
        bags[dup] = power_law_categories(rng, n_dup, n_cat, alpha)
    if permute:
        if perm is None:
            perm = rng.permutation(n_cat + 1)
        bags = perm[bags]
    if shuffle:
        keys = rng.random(bags.shape)
        order = np.argsort(keys, axis=1)
        bags = np.take_along_axis(bags, order, axis=1)
    elif permute:
        bags = np.sort(bags, axis=1)  # std::set order after permutation
    return bags


@dataclass
class Workload:
    """One synthetic batch, host arrays in the layout the API expects."""
    num_categories: int
    embed_width: int
    batch_size: int
    hotness: int
    indices: np.ndarray            # [nnz]
    offsets: Optional[np.ndarray]  # [batch+1] (CSR) or None
    weights: Optional[np.ndarray]  # [nnz] float32 values in {0.5, 0.25} or None
    nnz: int


def make_workload(num_categories: int, embed_width: int, batch_size: int,
                  hotness: int, alpha: float = 0.0, csr: bool = False,
                  weighted: bool = False, index_dtype=np.int32,
                  offset_dtype=np.int32, seed: int = 1234,
                  permute: bool = True, shuffle: bool = True) -> Workload:
    rng = np.random.default_rng(seed)
    bags = unique_bags(rng, batch_size, hotness, num_categories, alpha,
                       permute=permute, shuffle=shuffle)
    if csr:
        lens = rng.integers(0, hotness + 1, size=batch_size)
        offsets = np.zeros(batch_size + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        mask = np.arange(hotness)[None, :] < lens[:, None]
        indices = bags[mask]
        offsets = offsets.astype(offset_dtype)
    else:
        offsets = None
        indices = bags.reshape(-1)
    indices = np.ascontiguousarray(indices.astype(index_dtype))
    weights = None
    if weighted:
        weights = np.where(rng.random(indices.shape[0]) < 0.5, 0.5, 0.25)
        weights = weights.astype(np.float32)
    return Workload(num_categories, embed_width, batch_size, hotness, indices,
                    offsets, weights, int(indices.shape[0]))


def make_table(num_categories: int, embed_width: int, seed: int = 123456,
               dtype=np.float32) -> np.ndarray:
    """U(-1, 1) table, float32 host array (cast by the caller)."""
    rng = np.random.default_rng(seed)
    t = rng.random((num_categories, embed_width), dtype=np.float32)
    return (t * 2.0 - 1.0).astype(dtype)


def make_grad_y(rows: int, embed_width: int, seed: int = 654321) -> np.ndarray:
    """Integers in [-10, 10] as float32 (utils/src/embedding_allocation.cu:234-237)."""
    rng = np.random.default_rng(seed)
    return rng.integers(-10, 11, size=(rows, embed_width)).astype(np.float32)
