"""Enumerations of include/dsb200.h, mirrored for the Python host."""
ACT = {'identity': 0, 'relu': 1, 'elu': 2, 'tanh': 3, 'sigmoid': 4, 'leaky_relu': 5, 'softplus': 6, 'relu6': 7}
POOL = {'max': 0, 'average': 1}
STAT = {'mean': 0, 'var': 1, 'meanvar': 2, 'max': 3}
OPT_SGD, OPT_MOMENTUM, OPT_ADAM, OPT_RMSPROP = 0, 1, 2, 3
UNPOOL_REPEAT, UNPOOL_ZEROFILL = 0, 1
