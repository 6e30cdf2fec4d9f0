"""ampnet_b200: B200-native (sm_100a) implementation of the AMP-Net data-parallel hot path.

Directory name follows the repo convention (`3d-semantic-segmentation-amp-net_b200`); import it with
`importlib.import_module("3d-semantic-segmentation-amp-net_b200")` or through the root-level
shim `ampnet_b200.py`. Every op calls hand-written CUDA through the C ABI in
include/ampnet_b200.h; there is no CPU / PyTorch fallback.
"""
from . import _lib  # noqa: F401
from .sampling import fps, fps_batch, fps_indices, gather_rows, fps_host_batch, fps_host_stream, FpsHostStream  # noqa: F401
from .clustering import (kmeans_clustering, split_kmeans, split_kmeans_array, kmeans_assign,  # noqa: F401
                         kmeans_constrained_windows, regroup_windows, gather_feats, get_cluster_centroid)
from .modules import BasePointNet, TransformationNet, SegmentationWithAttention, set_default_precision  # noqa: F401
from .parallel import shard_windows, GradAllReduce  # noqa: F401
from .dataprep import split_windows, filter_normalize_windows  # noqa: F401
from .assembly import assemble_windows, draw_augmentation  # noqa: F401
from .blockstore import BlockStore, BlockStoreWriter, convert_kmeans_pt_files  # noqa: F401
from .tensorcore import tc_linear, linear_wgrad  # noqa: F401
from .graphstep import GraphedStep  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .streamed import StreamedForward  # noqa: F401

