from .analysis import network_to_rtd, pulse_moments, pulse_to_rtd, rtd_on_device  # noqa: F401
