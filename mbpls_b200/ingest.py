"""Host side of the ingest path: pandas DataFrames / pickled blocks -> page-locked host memory -> device.

The reference's loaders (mbpls/data/get_data.py:31-76) return dictionaries of pandas DataFrames read from pickles, and
``fit`` then copies every block three times on the host (deepcopy :303, check_array :310, StandardScaler :314) before the
``hstack`` (:379).  Here a block is staged ONCE, into page-locked memory in its stored precision (float32 stays float32),
so that ``MBPLS.fit`` streams it over PCIe at the copy engine's rate and the widening to float64, the feature-major
transposition and the standardisation all happen on the device (engine.ingest_feature_major, csrc/ingest.cu).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Sequence, Union

import numpy as np
import torch

__all__ = ["pin_block", "pin_blocks", "read_pickled_blocks"]


def pin_block(block, dtype=None) -> torch.Tensor:
    """One n x p_b array-like (numpy array, pandas DataFrame, nested list, CPU tensor) as a page-locked, row-major CPU
    tensor.  float32 and float64 sources keep their precision unless ``dtype`` says otherwise; everything else becomes
    float64 (what check_array(dtype=float64) produces, mbpls.py:310).  A tensor that is already pinned is returned as is."""
    if isinstance(block, torch.Tensor):
        if block.is_cuda:
            raise ValueError("pin_block stages HOST data; device tensors can be passed to fit directly")
        src = block
    else:
        arr = np.asarray(block)  # DataFrame -> its values (no copy for a single float block)
        if arr.ndim != 2:
            raise ValueError(f"Expected 2D array, got {arr.ndim}D array instead.")
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        if not arr.flags.writeable:
            arr = arr.copy() if not arr.flags.c_contiguous else arr
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            src = torch.from_numpy(np.ascontiguousarray(arr))
    want = src.dtype if dtype is None else dtype
    if src.dtype not in (torch.float32, torch.float64) and dtype is None:
        want = torch.float64
    if src.is_pinned() and src.dtype == want and src.is_contiguous():
        return src
    out = torch.empty(tuple(src.shape), dtype=want, pin_memory=True)
    out.copy_(src)
    return out


def pin_blocks(blocks: Union[Sequence, Dict], dtype=None) -> List[torch.Tensor]:
    """``pin_block`` for a list of blocks, or for the dictionary of DataFrames the reference's loaders return
    (values in insertion order)."""
    if isinstance(blocks, dict):
        blocks = list(blocks.values())
    return [pin_block(b, dtype) for b in blocks]


def read_pickled_blocks(paths: Iterable[str], dtype=None):
    """Pickled pandas DataFrames (the storage format of mbpls/data/**/*.pkl, read with ``pd.read_pickle`` at
    mbpls/data/get_data.py:39) -> (list of pinned blocks, list of names, list of column indexes).  Nothing is downloaded:
    a missing file raises FileNotFoundError (the reference falls back to a GitHub download, :88-100)."""
    import pandas as pd
    blocks, names, columns = [], [], []
    for path in paths:
        if not os.path.isfile(path):
            raise FileNotFoundError(path)
        df = pd.read_pickle(path)
        names.append(os.path.splitext(os.path.basename(path))[0])
        columns.append(getattr(df, "columns", None))
        blocks.append(pin_block(df, dtype))
    return blocks, names, columns
