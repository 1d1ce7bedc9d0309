"""B200-native endpoint picker for the LLM Instance Gateway ext-proc scheduler hot path.

Layout: ``csrc/`` holds the sm_100a kernels and the C ABI (include/lig.h, built to liblig.so);
the Python modules here are the host-side mirror of the reference's Go interfaces
(``backend`` record types, ``scheduling.Scheduler``) used by the tests and the benchmark.
"""
from .backend import Metrics, Pod, PodMetrics  # noqa: F401

__all__ = ["Pod", "Metrics", "PodMetrics"]
