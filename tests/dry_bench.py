"""bench.py's main() as a dry run on the CPU (test infrastructure; see tests/dry_gpu_plugin.py): the model bench.py builds
with the batch cut to B=2, T=64, cuda devices mapped to the CPU, CUDA graphs / events / streams mocked, kernels not executed
(validating dry library).  Timings and values are meaningless; what is exercised is every line of host code between the
argument parser and the JSON line.     python tests/dry_bench.py [bench.py flags]"""
import contextlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import torch  # noqa: E402
import dry_gpu_plugin as plugin  # noqa: E402

plugin.pytest_configure(None)
torch.cuda._sleep = lambda n: None
torch.cuda.is_current_stream_capturing = lambda: False
import torch.cuda.graphs as _tg  # noqa: E402
_tg.is_current_stream_capturing = lambda: False


class _Graph:
    def replay(self):
        pass


torch.cuda.CUDAGraph = _Graph
torch.cuda.graph = lambda g: contextlib.nullcontext()
_Adam = torch.optim.Adam
torch.optim.Adam = lambda params, **kw: _Adam(params, lr=kw.get("lr", 1e-3))     # fused / capturable are CUDA options

import bench  # noqa: E402

for k in bench.WORKLOADS:
    bench.WORKLOADS[k] = dict(bench.WORKLOADS[k], B=2, T=64, vocab=100)
bench.main()
