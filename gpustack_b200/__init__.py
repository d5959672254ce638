"""B200-native LLM inference engine for GPUStack's pluggable-engine surface.

Hot path: continuous-batched Llama-class prefill + decode on hand-written sm_100a kernels
(csrc/, C ABI in include/b200.h); boundary: `b200-engine serve` CLI + OpenAI-compatible HTTP
(cli.py, server.py).  There is no CPU execution path in this package.
"""
__version__ = "0.1.0"
