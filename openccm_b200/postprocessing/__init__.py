from .analysis import network_to_rtd, pulse_moments, pulse_to_rtd, rtd_on_device  # noqa: F401
from .writer import AsyncResultWriter, to_reference_layout  # noqa: F401
