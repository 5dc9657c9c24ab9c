"""stochastic_gcn_b200 -- B200-native (sm_100a) hot path of thu-ml/stochastic_gcn.

Neighbour sampler, sampled / full-neighbour aggregation (SpMM forward + backward) and the
control-variate history gather / write-back as hand-written CUDA kernels behind a C ABI
(include/sgcn_b200.h, stochastic_gcn_b200/libsgcn_b200.so), with Python mirrors of the reference's
operator surfaces:

  scheduler.PyScheduler              <- gcn/_scheduler.pyx   (ext module `scheduler`)
  history.slice / dense_slice        <- gcn/_history.pyx     (ext module `history`)
  layers.PlainAggregator/VRAggregator <- gcn/layers.py:214-362

There is no CPU implementation in this package: every entry point raises if the CUDA library is
missing or a kernel launch fails.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "sampler", "scheduler", "history", "layers", "graphs", "step", "sharding"]
__version__ = "0.1.0"

import os as _os

# The step drivers run five to six CUDA streams side by side, some of whose kernels wait on the device for
# others (bounded spins).  With the default of 8 hardware work queues streams share queues, and a queue
# processes its entries in order -- a waiting kernel's stream can then hold back an unrelated one.  One queue
# per stream avoids that; it only takes effect if set before the CUDA context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# Lazy module loading (the CUDA default) can make the FIRST launch of a kernel wait for the device to go idle --
# behind a thread block that waits for that very kernel.  The step drivers therefore run their first pass without
# device-side waits on later submissions (csrc/step.cu); CUDA_MODULE_LOADING=EAGER removes the hazard for direct
# users of the fused / train entry points too, at the price of loading every kernel of every CUDA library in the
# process at start-up (tens of seconds on a cold machine) -- not set here.
