"""
Stand-alone stand-in for the slice of OpenCCM's ``ConfigParser`` that the solver path reads.

The device solvers accept *any* object exposing ``get_item``, ``get_list``, ``get_expression`` and
``parser[section][key]`` — in production that is ``openccm.ConfigParser``
(/root/reference/openccm/config_functions/expanded_config_parser.py:75-326).  This class exists so the
path can be exercised where the reference package is absent (the GPU test box).  It implements the same
accessor semantics (``', '`` list separator :260, ``== 'True'`` booleans :267, defaults of the SIMULATION /
POST-PROCESSING sections :52-67, ``points_per_pfr`` forced to 1 for CSTR :110-115) and nothing else
(no path rewriting, no OpenCMP/OpenFOAM validation).
"""
from __future__ import annotations

import ast
import configparser
from typing import Dict, Optional

HOT_PATH_DEFAULTS: Dict[str, Dict[str, str]] = {
    'SETUP': {'tmp_folder_path': 'cache/', 'output_folder_path': 'output_ccm/', 'working_directory': './'},
    'INPUT': {},
    'COMPARTMENT MODELLING': {},
    'SIMULATION': {'run': 'True', 'points_per_pfr': '2', 'first_timestep': '0.0001', 'solver': 'LSODA',
                   'reactions_file_path': 'None', 'rtol': '1e-06', 'atol': '1e-06', 't_eval': 'all',
                   'boundary_conditions': ''},
    'POST-PROCESSING': {'calculate_rtd': 'False'},
}


class ConfigParser(configparser.ConfigParser):
    def __init__(self, path: Optional[str] = None, *, text: Optional[str] = None):
        super().__init__()
        self.need_to_update_paths = False
        if path is not None:
            with open(path) as fh:
                self.read_string(fh.read())
        elif text is not None:
            self.read_string(text)
        model = self.get_item(['COMPARTMENT MODELLING', 'model'], str).lower()
        if model not in ('cstr', 'pfr'):
            raise ValueError(f"Invalid model type ({model}) specified.")
        self['COMPARTMENT MODELLING']['model'] = model
        if model == 'cstr':
            if not self.has_section('SIMULATION'):
                self.add_section('SIMULATION')
            self['SIMULATION']['points_per_pfr'] = '1'
        for section, entries in HOT_PATH_DEFAULTS.items():
            if not self.has_section(section):
                self.add_section(section)
            for key, val in entries.items():
                if key not in self[section]:
                    self[section][key] = val

    def _raw(self, keys):
        section, key = keys
        try:
            return self[section][key]
        except KeyError:
            raise ValueError(f"Need to specify a value for {section}, {key}")

    def get_item(self, keys, val_type):
        raw = self._raw(keys)
        return raw.lower() == 'true' if val_type is bool else val_type(raw)

    def get_list(self, keys, val_type):
        parts = self._raw(keys).split(', ')
        return [(p == 'True') if val_type is bool else val_type(p) for p in parts]

    def get_expression(self, keys):
        return ast.literal_eval(self._raw(keys))

    def __hash__(self):
        return hash(tuple(hash(v) for s in self.sections() for v in self[s].values()))


class GroupedBCs:
    """Boundary-name -> negative integer id, as `openccm.mesh.GroupedBCs.id` (mesh/cmesh.py:91-115):
    sorted inlets then sorted outlets are -1, -2, ...; ignored names follow; all no-flux names share one id."""

    def __init__(self, config_parser):
        def names(key):
            try:
                v = config_parser.get_expression(['INPUT', key])
            except (ValueError, KeyError):
                v = None
            return tuple(v) if v else tuple()
        self.domain_inlet_names = tuple(sorted(names('domain_inlet_names')))
        self.domain_outlet_names = tuple(sorted(names('domain_outlet_names')))
        self.domain_in_out_names = self.domain_inlet_names + self.domain_outlet_names
        self.ignored_names = names('ignored_names')
        self.no_flux_names = names('no_flux_names')
        self.no_flux = -(1 + len(self.domain_in_out_names) + len(self.ignored_names))
        self.domain_inlets = tuple(self.id(n) for n in self.domain_inlet_names)
        self.domain_outlets = tuple(self.id(n) for n in self.domain_outlet_names)

    def id(self, label: str) -> int:
        if label in self.no_flux_names:
            return self.no_flux
        if label in self.domain_in_out_names:
            return -(1 + self.domain_in_out_names.index(label))
        if label in self.ignored_names:
            return -(1 + len(self.domain_in_out_names) + self.ignored_names.index(label))
        raise ValueError(f"The passed label ({label}) is not one of the specified bcs: "
                         f"{self.no_flux_names + self.domain_in_out_names + self.ignored_names}")
