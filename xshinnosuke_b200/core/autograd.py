"""Reverse-mode walk over the dynamic graph (mirror of xs/core/autograd.py:5-45).

Same contract as the reference: depth-first topological sort from the loss through `in_bounds`
(Parameters and None are not descended into) down to GLOBAL.INPUTS; closures are called in reverse
order as `node.grad_fn(node)` and accumulate into `in_bound.grad`; a non-leaf node's activation and
gradient are released right after its closure ran (returned to the caching device allocator) unless
`retain_graph` / `retain_grad`. The data-parallel engine hooks GLOBAL.AFTER_NODE_BACKWARD to launch
gradient-bucket all-reduces while the remaining closures still run.
"""
import sys

from . import state as GLOBAL
from .tensor import Parameter


def backward(outputs, inputs, retain_graph=False):
    graph = topological_sort(inputs, outputs)
    hook = GLOBAL.AFTER_NODE_BACKWARD
    for node in reversed(graph):
        if node.grad_fn is not None:
            node.grad_fn(node)
            if hook is not None:
                hook(node)
        for layer in node.next_layers:
            layer.data = None
        if retain_graph or node.is_leaf or node.retain_grad():
            node.reset_(requires_grad=node.requires_grad, name=node.name, slices=node.slices(),
                        static_graph_tensor=node.is_static, retain_grad=node.retain_grad(), grad=node.grad)
        else:
            node.free_memory()
            node.reset_(requires_grad=node.requires_grad, name=node.name, slices=node.slices(),
                        static_graph_tensor=node.is_static)
    if retain_graph:
        GLOBAL.GRAPH = graph
    else:
        del graph[:]
        GLOBAL.INPUTS = None
        GLOBAL.OUTPUTS = None
        GLOBAL.GRAPH = None


def topological_sort(ins, outs):
    """Post-order DFS (iterative: ResNet graphs are deeper than the default recursion limit allows
    once Python frames of closures are added)."""
    topo, visited = [], set()
    stack = [(outs, iter(()), False)]
    # emulate the reference's recursion exactly: a node is appended after all its children
    stack = [(outs, None)]
    while stack:
        v, it = stack.pop()
        if it is None:
            if ins is not None and v is ins:
                topo.append(v)
                continue
            if id(v) in visited:
                continue
            visited.add(id(v))
            it = iter(v.in_bounds)
        advanced = False
        for child in it:
            if child is None or isinstance(child, Parameter):
                continue
            stack.append((v, it))
            stack.append((child, None))
            advanced = True
            break
        if not advanced:
            topo.append(v)
    return topo


def reset_graph(graph):
    del graph[:]
    GLOBAL.INPUTS = None
    GLOBAL.OUTPUTS = None
    GLOBAL.GRAPH = None
