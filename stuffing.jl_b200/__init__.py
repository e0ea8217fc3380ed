"""stuffing.jl_b200 — B200-native collision-and-fit engine behind the interface of Stuffing.jl.

Only what the hot path needs: csrc/ (sm_100a kernels + C ABI, built into libstuffing_b200.so) and
the host-side mirror of the reference's QTrees / Trainer interface.  The directory name is not a
Python identifier; load it with `import _pkg; S = _pkg.load()` (repo root) — it registers itself
as module `stuffing_b200`.
"""
from . import _lib  # noqa: F401
from .qtrees import *  # noqa: F401,F403
from .qtrees import (DeviceQTrees, DeviceTree, HostTree, DynamicColliders, EMPTY, FULL, MIX, collision, collision_dfs,
                     collision_randbfs, collision_bfs_items, dynamiccollisions, getpositions, hash_spacial_qtree,
                     items_to_array, items_to_list, linked_spacial_qtree, locate, maskqtree, partialcollisions, qtree, qtrees,
                     setpositions, totalcollisions, totalcollisions_native, totalcollisions_native_items, totalcollisions_spacial)
from . import trainer  # noqa: F401
from .trainer import Momentum, SGD, batchsteps, fit, fit_round, grad2d, train  # noqa: F401
from . import place as placement  # noqa: F401
from .place import place, overlap_ground  # noqa: F401
from . import io  # noqa: F401,E402
from .io import charimage, charmat, load_scene, packing, save_scene  # noqa: F401,E402
