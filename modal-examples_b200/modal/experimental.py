"""``modal.experimental`` names used in the tree: ``http_server``, ``clustered`` / ``get_cluster_info``
(14_clusters/simple_torch_cluster.py:96-130).  In-box a "cluster" is the single 8-GPU host: rank 0 of 1."""
from dataclasses import dataclass, field

from .web import _inert

http_server = _inert("http_server")


def clustered(size: int = 1, **_kw):
    from .cls import _mark

    return lambda fn: _mark(fn, clustered=size)


@dataclass
class ClusterInfo:
    rank: int = 0
    container_ips: list = field(default_factory=lambda: ["127.0.0.1"])
    container_ipv4_ips: list = field(default_factory=lambda: ["127.0.0.1"])
    cluster_id: str = "in-box"


def get_cluster_info() -> ClusterInfo:
    return ClusterInfo()


def stop_fetching_inputs():
    return None
