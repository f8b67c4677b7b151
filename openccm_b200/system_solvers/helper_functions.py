"""Output-time generation, mirror of /root/reference/openccm/system_solvers/helper_functions.py:56-110."""
from __future__ import annotations

from typing import List, Optional

import numpy as np


def generate_t_eval(config_parser) -> Optional[List[float]]:
    """``t_eval`` option -> list of output times (None = every solver step).

    Forms: ``all`` | ``dt, linear`` | ``log, n`` | ``t1, t2, ..., tn``.  The linear form accumulates ``t += dt`` like
    the reference (:89-94) so that the produced doubles are identical, and always ends with ``tf``."""
    spec = config_parser.get_list(['SIMULATION', 't_eval'], str)
    t0, tf = config_parser.get_list(['SIMULATION', 't_span'], float)
    assert t0 >= 0
    if len(spec) == 1 and spec[0].lower() == 'all':
        return None
    if len(spec) == 2 and 'linear' in spec[1].lower():
        dt = float(spec[0])
        out, t = [], t0
        while t < tf:
            out.append(t)
            t += dt
        out.append(tf)
        return out
    if len(spec) == 2 and 'log' in spec[0].lower():
        n = int(spec[1])
        first = config_parser.get_item(['SIMULATION', 'first_timestep'], float)
        if t0 == 0:
            ts = np.zeros(n)
            ts[1:] = np.logspace(np.log10(first), np.log10(tf), n - 1)
        else:
            ts = np.logspace(np.log10(t0), np.log10(tf), n)
        ts[0], ts[-1] = t0, tf
        return [float(v) for v in ts]
    out = sorted(float(v) for v in spec)
    if out[0] < t0 or out[-1] > tf:
        raise ValueError('Invalid time range provided, times must be within range specified by t_range.')
    return out
