"""Drop-in for ``wsovod.layers`` (wsovod/layers/__init__.py:1-2): same names, same call signatures."""
from .roi_loop_pool import ROIAlign, ROILoopPool, RoIPool, roi_loop_pool
from .csc import CSC, CSCConstraint, csc, csc_constraint

__all__ = ["ROILoopPool", "roi_loop_pool", "RoIPool", "ROIAlign", "CSC", "CSCConstraint", "csc", "csc_constraint"]
