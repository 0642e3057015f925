"""Constants, column naming and the progress wrapper.

Mirrors /root/reference/pysrc/cityseer/config.py (constants :60-64, prep_gdf_key :21-32, wrap_progress :96-130).
"""
from __future__ import annotations

import logging
import os
import threading
import time
from collections.abc import Callable
from queue import Queue
from typing import Any

import numpy as np

logger = logging.getLogger(__name__)
np.seterr(invalid="ignore")


def prep_gdf_key(key: str, dist: int, angular: bool = False, weighted: bool | None = None) -> str:
    """Format a result column label: ``cc_{key}_{dist}[_ang][_wt|_nw]``."""
    key = key.replace(".0", "")
    key = key.replace(".0_", "_")
    key = f"cc_{key}_{dist}"
    if angular is True:
        key += "_ang"
    if weighted is True:
        key += "_wt"
    elif weighted is False:
        key += "_nw"
    return key


def check_quiet() -> bool:
    if "GCP_PROJECT" in os.environ:
        return True
    return os.environ.get("CITYSEER_QUIET_MODE", "").lower() in ["true", "1"]


QUIET_MODE = check_quiet()
DEBUG_MODE: bool = os.environ.get("CITYSEER_DEBUG_MODE", "").lower() in ["true", "1"]
SKIP_VALIDATION: bool = False
MIN_THRESH_WT: float = 0.01831563888873418
SPEED_M_S: float = 1.33333
ATOL: float = 0.01
RTOL: float = 0.0001


def log_thresholds(distances: list[int], betas: list[float], seconds: list[int]) -> None:
    logger.info("Metrics computed for:")
    for d, b, s in zip(distances, betas, seconds):
        logger.info(f"Distance: {d}m, Beta: {round(b, 5)}, Walking Time: {s / 60} minutes.")


def wrap_progress(total: int, rust_struct: Any, partial_func: Callable, desc: str | None = None) -> Any:
    """Run ``partial_func`` on a worker thread while the caller polls ``rust_struct.progress()`` at 10 Hz.

    The native call releases the GIL (ctypes does so for every foreign call), so polling works exactly as with
    the reference's PyO3 ``py.detach``.
    """
    try:
        from tqdm import tqdm
    except Exception:  # pragma: no cover - tqdm is present in this image
        tqdm = None

    def wrapper(queue: Queue):
        try:
            queue.put(partial_func())
        except Exception as e:  # noqa: BLE001 - re-raised in the caller, as the reference does
            queue.put(e)

    result_queue: Queue = Queue()
    thread = threading.Thread(target=wrapper, args=(result_queue,))
    pbar = tqdm(total=total, disable=QUIET_MODE, desc=desc) if tqdm is not None else None
    thread.start()
    while thread.is_alive():
        time.sleep(0.1)
        if pbar is not None:
            pbar.update(rust_struct.progress() - pbar.n)
    if pbar is not None:
        pbar.update(total - pbar.n)
        pbar.close()
    result = result_queue.get()
    thread.join()
    if isinstance(result, Exception):
        raise result
    return result
