"""
Asynchronous result writer (second half): results leave the GPU and reach the disk while the device keeps working.

The reference writes its results synchronously at the end of a run: ``np.save(<model>_concentrations.npy)`` / ``_t.npy``
(run.py:200-201, layout ``(S, n_points, T)``), and one text file per (time, species) in ``export_to_vtu_openfoam``
(postprocessing/vtu_output.py:311-322).  Here a device tensor is copied to a pinned host buffer on a dedicated copy stream
(ordered after the producing kernels by an event), and a worker thread waits for the copy, applies an optional host-side
transform (e.g. the permutation from the device layout ``[T, n_points, S]`` to the reference's ``(S, n_points, T)``) and writes
the file — the solver stream never blocks, and at most ``max_pending`` buffers are in flight.  Host arrays take the same path
without the copy.
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Optional, Union

import numpy as np
import torch


def to_reference_layout(n_points: int, n_species: int) -> Callable[[np.ndarray], np.ndarray]:
    """[T, n_points * S] (device layout, as recorded by the integrators) -> (S, n_points, T) (cstr_system.py:152-154)."""
    def f(y: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(y.reshape(y.shape[0], n_points, n_species).transpose(2, 1, 0))
    return f


class AsyncResultWriter:
    def __init__(self, max_pending: int = 2):
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, int(max_pending)))
        self._err: Optional[BaseException] = None
        self._copy_stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self._thread = threading.Thread(target=self._work, daemon=True)
        self._thread.start()
        self.n_written = 0

    # ------------------------------------------------------------------ producer side
    def submit(self, sink: Union[str, Callable[[np.ndarray], None]], data: Union[torch.Tensor, np.ndarray],
               transform: Optional[Callable[[np.ndarray], np.ndarray]] = None) -> None:
        """Queue ``data`` for ``np.save(sink, ...)`` (or ``sink(array)`` when callable). Blocks only when ``max_pending`` are in flight."""
        if self._err is not None:
            raise self._err
        if isinstance(data, torch.Tensor) and data.is_cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(data.device))          # after the kernels that produced `data`
            host = torch.empty(data.shape, dtype=data.dtype, pin_memory=True)
            done = torch.cuda.Event()
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ready)
                host.copy_(data, non_blocking=True)
                done.record(self._copy_stream)
            data.record_stream(self._copy_stream)                          # the allocator must not recycle it before the copy ran
            self._q.put((sink, host, done, transform))
        else:
            arr = data.numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
            self._q.put((sink, arr, None, transform))

    def close(self) -> int:
        """Wait for everything queued; returns the number of arrays written. Re-raises a worker failure."""
        self._q.put(None)
        self._thread.join()
        if self._err is not None:
            raise self._err
        return self.n_written

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ worker
    def _work(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            sink, host, done, transform = item
            try:
                if done is not None:
                    done.synchronize()
                    host = host.numpy()
                arr = transform(host) if transform is not None else host
                if callable(sink):
                    sink(arr)
                else:
                    np.save(sink, arr)
                self.n_written += 1
            except BaseException as exc:          # surfaced by the next submit() / close()
                self._err = exc
