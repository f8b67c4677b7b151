from .cstr import connect_cstr_compartments, create_cstr_network, create_cstr_network_arrays  # noqa: F401
from .helpers import tweak_final_flows  # noqa: F401
