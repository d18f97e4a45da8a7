"""Parallelism: data parallel engines (reference-semantics and native) and SyncBN."""
from .syncbn import SyncBatchNorm  # noqa: F401
from .torch_ddp import BucketedDataParallel, plan_buckets  # noqa: F401
