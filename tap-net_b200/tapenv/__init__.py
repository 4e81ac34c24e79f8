"""tapenv -- B200-native packing-environment step for TAP-Net's decode loop.

Public surface (reference names, reference signatures):
    update_dynamic, update_mask          pack.py:333-376, :276-331
    Container                            tools.py:3607-3966  (per-environment view)
    BatchedContainers                    the batch of containers model.py:294 builds, as one HBM state buffer
The kernels live in lib/libtapenv.so (csrc/, sm_100a) behind the C ABI of include/tapenv.h; importing this
package without that library raises -- there is no CPU fallback.
"""
from . import _capi
from ._capi import TapEnvError
from .config import make_config, rotate_types
from .ops import update_dynamic, update_mask
from .containers import BatchedContainers, BatchedContainerPairs, Container
from .runner import EpisodeRunner, HostPipeline
from .decode import DecodeLoop, RollingDecodeLoop
from . import adapters, dist, generators
from .dataset import PACKDataset, pack_inputs
from .episode import calc_positions_lb_greedy, calc_positions_mcs, reward
from .dropin import install, uninstall
from .rolling import BatchedInitialContainers, RollingRunner, RollingHostPipeline, pack_graphs

__all__ = ["PACKDataset", "pack_inputs", "reward", "calc_positions_lb_greedy", "calc_positions_mcs", "install", "uninstall",
           "update_dynamic", "update_mask", "Container", "BatchedContainers", "BatchedContainerPairs", "EpisodeRunner", "DecodeLoop", "RollingDecodeLoop", "HostPipeline", "make_config", "rotate_types",
           "TapEnvError", "BatchedInitialContainers", "RollingRunner", "RollingHostPipeline", "pack_graphs"]
