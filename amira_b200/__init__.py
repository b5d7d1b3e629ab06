"""amira_b200 -- B200-native (sm_100a CUDA) gene-space de Bruijn graph build for Amira.

Public surface mirrors the upstream modules on the graph-build path:
``GeneMerGraph`` (amira/construct_graph.py), ``build_graph`` / ``build_multiprocessed_graph``
(amira/graph_utils.py), and the element classes ``Gene``, ``GeneMer``, ``Read``, ``Node``, ``Edge``.
``DeviceGraph`` is the array-level object over the C ABI (include/amira_gmg.h)."""
from .construct_edge import Edge
from .construct_gene import Gene, hashlib_hash
from .construct_gene_mer import GeneMer
from .construct_graph import GeneMerGraph, bind_upstream
from .construct_node import Node
from .construct_read import Read
from .device_graph import DeviceGraph
from .encode import EncodedReads
from .graph_utils import build_graph, build_multiprocessed_graph, get_overall_mean_node_coverages

__all__ = ["GeneMerGraph", "bind_upstream", "build_graph", "build_multiprocessed_graph", "DeviceGraph", "EncodedReads", "get_overall_mean_node_coverages",
           "Gene", "GeneMer", "Read", "Node", "Edge", "hashlib_hash"]
__version__ = "0.1.0"
