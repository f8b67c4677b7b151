from .helpers import tweak_final_flows  # noqa: F401
