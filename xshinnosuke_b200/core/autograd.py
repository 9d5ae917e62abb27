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


def _recycle(node, keep):
    """Return a visited node to its pre-forward state. `keep` (leaf, retain_grad or retain_graph): payload and gradient
    survive; otherwise both go back to the caching device allocator right here, inside the backward walk."""
    state = dict(requires_grad=node.requires_grad, name=node.name, slices=node.slices(), static_graph_tensor=node.is_static)
    if keep:
        state.update(retain_grad=node.retain_grad(), grad=node.grad)
    else:
        node.free_memory()
    node.reset_(**state)


def backward(outputs, inputs, retain_graph=False):
    order = topological_sort(inputs, outputs)
    if GLOBAL.BEFORE_BACKWARD is not None:
        GLOBAL.BEFORE_BACKWARD()
    after_node = GLOBAL.AFTER_NODE_BACKWARD
    for node in reversed(order):                       # loss first, inputs last
        closure = node.grad_fn
        if closure is not None:
            closure(node)                              # accumulates into the gradients of node.in_bounds
            if after_node is not None:
                after_node(node)
        for layer in node.next_layers:                 # layers that consumed this tensor drop their handle on it
            layer.data = None
        _recycle(node, keep=retain_graph or node.is_leaf or node.retain_grad())
    if retain_graph:
        GLOBAL.GRAPH = order
    else:
        reset_graph(order)


def topological_sort(ins, outs):
    """Post-order depth-first walk from the loss through `in_bounds` (a node is listed after everything it was computed
    from; Parameters and None are not descended into; the walk stops at `ins`). Iterative with an explicit stack of
    (node, iterator over its operands): ResNet graphs are deeper than Python's recursion limit allows."""
    order, seen = [], set()
    stack = [(outs, None)]
    while stack:
        node, operands = stack.pop()
        if operands is None:                           # first visit
            if ins is not None and node is ins:
                order.append(node)
                continue
            if id(node) in seen:
                continue
            seen.add(id(node))
            operands = iter(node.in_bounds)
        child = next((c for c in operands if c is not None and not isinstance(c, Parameter)), None)
        if child is None:
            order.append(node)                         # all operands done
        else:
            stack.append((node, operands))             # come back for the remaining operands
            stack.append((child, None))
    return order


def reset_graph(graph):
    del graph[:]
    GLOBAL.INPUTS = None
    GLOBAL.OUTPUTS = None
    GLOBAL.GRAPH = None
