"""
Results -> mesh elements.

Mirror of ``openccm.postprocessing.vtu_output.export_to_vtu_openfoam`` (vtu_output.py:279-336): the reference fills one
buffer per (time, species) with a Python loop over every DOF's element list; here the interpolation table is built once,
vectorised (``export.element_interpolation``), and all fields come out of one kernel launch (``occm_dofs_to_elements``).
The OpenFOAM file formatting itself is not part of the hot path: ``export_to_vtu_openfoam`` hands the buffers to the
reference's own ``write_buffer_to_file_openfoam`` when ``openccm`` is importable, or to the ``write_buffer`` callable given.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..system_solvers.export import element_interpolation


def element_fields(concentrations_all_time, model_to_element_map: Optional[list], points_per_model: int, n_elements: int,
                   *, device_layout: bool = False, as_numpy: bool = True):
    """fields[T, S, n_elements] of ``c[S, n_points, T]`` (the array ``solve_system`` returns), -1 outside the compartments.

    ``device_layout=True``: the input is a CUDA tensor [T, n_points * S] as recorded by the integrators (no host round trip)."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError('openccm_b200 needs a CUDA device (no CPU fallback)')
    dev = torch.device('cuda', torch.cuda.current_device())
    if device_layout:
        y = concentrations_all_time
        assert y.is_cuda and y.dtype == torch.float64 and y.dim() == 2 and y.is_contiguous()
        T = y.shape[0]
        n_points = (len(model_to_element_map) if model_to_element_map is not None else None)
        n_points = n_points * points_per_model if n_points is not None else None
        S = y.shape[1] // n_points
    else:
        c = np.asarray(concentrations_all_time, dtype=np.float64)
        S, n_points, T = c.shape
        y = torch.as_tensor(np.ascontiguousarray(c.transpose(2, 1, 0)).reshape(T, n_points * S), device=dev)
    n_models = n_points // points_per_model
    el_dof, el_other, el_w = element_interpolation(model_to_element_map, n_models, points_per_model, n_elements)
    d_dof = torch.as_tensor(el_dof, device=dev); d_other = torch.as_tensor(el_other, device=dev); d_w = torch.as_tensor(el_w, device=dev)
    out = torch.empty((T, S, n_elements), dtype=torch.float64, device=dev)
    _lib.check(lib.occm_dofs_to_elements(y.data_ptr(), T, S, n_points * S, n_elements, d_dof.data_ptr(), d_other.data_ptr(),
                                         d_w.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out.cpu().numpy() if as_numpy else out


def export_to_vtu_openfoam(concentrations_all_time: np.ndarray, ts: np.ndarray,
                           model_to_element_map: Optional[List[List[Tuple[float, int]]]], config_parser, cmesh,
                           dof_to_element_map=None, *, write_buffer: Optional[Callable] = None) -> None:
    """Same signature, folder layout and files as the reference (vtu_output.py:279-336)."""
    print("Exporting simulation visualization")
    if dof_to_element_map:
        raise NotImplementedError('pass model_to_element_map; the interpolation table is rebuilt vectorised')
    if write_buffer is None:
        from openccm.postprocessing.vtu_output import write_buffer_to_file_openfoam as write_buffer   # the reference's formatter
    output_folder_path = config_parser.get_item(['SETUP', 'output_folder_path'], str)
    vtu_folder_path = config_parser.get_item(['POST-PROCESSING', 'vtu_dir'], str)
    points_per_model = config_parser.get_item(['SIMULATION', 'points_per_pfr'], int)
    species_names = config_parser.get_list(['SIMULATION', 'specie_names'], str)
    t_span = config_parser.get_list(['SIMULATION', 't_span'], float)
    num_elements = len(cmesh.element_sizes)
    assert len(model_to_element_map) * points_per_model == np.shape(concentrations_all_time)[1]
    # all (time, species) element buffers in one kernel launch; they leave the device one output time at a time through the
    # asynchronous writer, so the (slow, text) file formatting of time i overlaps the download of time i + 1
    from .writer import AsyncResultWriter
    fields = element_fields(concentrations_all_time, model_to_element_map, points_per_model, num_elements, as_numpy=False)

    t0_path = os.path.join(output_folder_path + vtu_folder_path + str(t_span[0]))
    contents_of_t0 = [(name, os.path.abspath(os.path.join(t0_path, name))) for name in os.listdir(t0_path)]

    def sink_for(i_t, t):
        def sink(buf):                                          # buf: [S, n_elements] of this output time, on the host
            output_folder = output_folder_path + vtu_folder_path + str(t)
            Path(output_folder).mkdir(parents=True, exist_ok=(i_t == 0))
            if t != t_span[0]:
                for name, original in contents_of_t0:
                    os.symlink(original, os.path.join(output_folder, name))
            for i_specie, specie in enumerate(species_names):
                write_buffer(buf[i_specie], output_folder + '/c_' + specie, t, specie, cmesh)
        return sink

    with AsyncResultWriter(max_pending=2) as writer:
        for i_t, t in enumerate(ts):
            writer.submit(sink_for(i_t, t), fields[i_t])
    print("Done exporting simulation visualization")
