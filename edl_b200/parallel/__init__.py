"""Data-parallel engine: flat parameter storage, NVSwitch symmetric memory, fused all-reduce
planning / overlap, elastic re-planning."""
from .flat import FlatParams
from .ddp import ElasticDataParallel, plan_buckets, choose_algo
from .dgc import DGCMomentum

__all__ = ["FlatParams", "ElasticDataParallel", "plan_buckets", "choose_algo", "DGCMomentum"]
