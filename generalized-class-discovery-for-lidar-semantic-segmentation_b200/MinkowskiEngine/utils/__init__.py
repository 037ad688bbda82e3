from gcdlss_b200.nn import kaiming_normal_
from gcdlss_b200.quantize import batched_coordinates, sparse_quantize

__all__ = ["sparse_quantize", "batched_coordinates", "kaiming_normal_"]
