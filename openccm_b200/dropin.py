"""
Drop-in installation behind OpenCCM's own entry point.

``openccm.run`` binds ``solve_system`` at import time (``from .system_solvers import solve_system``,
/root/reference/openccm/run.py:35) and calls it at run.py:199, so both names must be rebound
(SURVEY.md §8(b)).  After ``install()`` an unmodified ``openccm`` run uses the device whenever the CONFIG says

    [SIMULATION]
    solver = B200-RK45        # or B200-BDF

``install()`` also rebinds ``network_to_rtd`` (run.py:205) to the device RTD kernel; ``uninstall()`` restores both.

Every other solver name is forwarded untouched to the reference's scipy path (this package itself never integrates on
the CPU).  With ``route_scipy_names=True`` the scipy method names are mapped too: RK45/RK23/DOP853 -> B200-RK45,
LSODA -> B200-LSODA (RK45 probe, then BDF if the step is stability limited), BDF/Radau -> B200-BDF.
"""
from __future__ import annotations

from typing import Callable, Optional

from .system_solvers import is_device_solver, solve_system as _device_solve_system

_SCIPY_TO_DEVICE = {'RK45': 'B200-RK45', 'RK23': 'B200-RK45', 'DOP853': 'B200-RK45',
                    'LSODA': 'B200-LSODA', 'BDF': 'B200-BDF', 'RADAU': 'B200-BDF'}
_original: Optional[Callable] = None
_original_rtd: Optional[Callable] = None
# every name `network_to_rtd` is bound to inside the reference (run.py:32-33 imports it from the package, which imports it
# from .analysis, postprocessing/__init__.py:28)
_RTD_MODULES = ('openccm.postprocessing.analysis', 'openccm.postprocessing', 'openccm.run')


def make_dispatcher(reference_solve_system: Callable, route_scipy_names: bool = False) -> Callable:
    def solve_system(model, model_network, config_parser, cmesh):
        solver = config_parser.get_item(['SIMULATION', 'solver'], str)
        if route_scipy_names and solver.upper() in _SCIPY_TO_DEVICE:
            config_parser['SIMULATION']['solver'] = _SCIPY_TO_DEVICE[solver.upper()]
            solver = config_parser['SIMULATION']['solver']
        if is_device_solver(solver):
            return _device_solve_system(model, model_network, config_parser, cmesh)
        return reference_solve_system(model, model_network, config_parser, cmesh)
    solve_system.__doc__ = reference_solve_system.__doc__
    return solve_system


def install(route_scipy_names: bool = False, rtd: bool = True) -> None:
    """Rebind openccm.system_solvers.solve_system and openccm.run.solve_system; with ``rtd`` also every binding of
    ``network_to_rtd`` (run.py:205 then computes the residence-time distribution with occm_rtd on the device).
    Needs the `openccm` package importable."""
    global _original, _original_rtd
    import importlib
    ss = importlib.import_module('openccm.system_solvers')
    # `openccm/__init__.py` re-exports the function `run`, which shadows the submodule attribute: take the module itself
    run_mod = importlib.import_module('openccm.run')
    if _original is None:
        _original = ss.solve_system
    dispatcher = make_dispatcher(_original, route_scipy_names)
    ss.solve_system = dispatcher
    run_mod.solve_system = dispatcher
    if rtd:
        from .postprocessing.analysis import network_to_rtd as device_rtd
        mods = [importlib.import_module(name) for name in _RTD_MODULES]
        if _original_rtd is None:
            _original_rtd = mods[0].network_to_rtd

        def network_to_rtd(system_results, c_mesh, config_parser, network):
            return device_rtd(system_results, c_mesh, config_parser, network)
        network_to_rtd.__doc__ = _original_rtd.__doc__
        for mod in mods:
            mod.network_to_rtd = network_to_rtd


def uninstall() -> None:
    global _original, _original_rtd
    import importlib
    if _original is not None:
        ss = importlib.import_module('openccm.system_solvers')
        run_mod = importlib.import_module('openccm.run')
        ss.solve_system = _original
        run_mod.solve_system = _original
        _original = None
    if _original_rtd is not None:
        for name in _RTD_MODULES:
            importlib.import_module(name).network_to_rtd = _original_rtd
        _original_rtd = None
