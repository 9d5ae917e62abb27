"""Process-global graph / mode state (the reference's xs/core/__global.py:1-9).

INPUTS / OUTPUTS / GRAPH delimit the dynamic graph of the step in flight; TRAINING and
COMPUTE_GRAD gate BN statistics and grad_fn wiring. One model step at a time per process.
"""
INPUTS = None
OUTPUTS = None
GRAPH = None
TRAINING = True
IS_TRAINING = True      # written by train()/eval(); like the reference, ops read TRAINING
COMPUTE_GRAD = True
USE_CUDA = True         # always: this backend has no host compute path

# data-parallel hooks: called by core.autograd.backward / Model.backward before the first closure of a backward pass
# and after every node's closure ran
BEFORE_BACKWARD = None
AFTER_NODE_BACKWARD = None
