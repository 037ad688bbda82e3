from gcdlss_b200.nn import BasicBlock, Bottleneck

__all__ = ["BasicBlock", "Bottleneck"]
