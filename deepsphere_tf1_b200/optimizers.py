"""Optimizer / schedule descriptions with the names the reference's experiment drivers use.

The reference passes `scheduler = lambda step: tf.train.exponential_decay(...)` and
`optimizer = lambda lr: tf.train.AdamOptimizer(lr, epsilon=1e-8)` to the model
(experiments/cosmo/experiment_deepsphere.py:97-98, experiments/shrec17/hyperparameters.py:146-150)
and `base_model.training` applies them (deepsphere/models.py:822-843).  Here the same two
lambdas are written against this module (`from deepsphere_tf1_b200 import optimizers as train`):
`scheduler(step) -> float`, `optimizer(lr) -> one of the classes below`.  The update rules are
TF 1.10's (Adam's epsilon sits outside the bias correction; RMSProp's `ms` slot starts at 1)
and run as one fused kernel over the flat parameter buffer (dsb200_optimizer_step).
"""
from ._consts import OPT_ADAM, OPT_MOMENTUM, OPT_RMSPROP, OPT_SGD


def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
    """tf.train.exponential_decay: lr * rate^(step / decay_steps)."""
    e = float(global_step) / float(decay_steps)
    if staircase:
        e = float(int(e))
    return float(learning_rate) * float(decay_rate) ** e


class _Optimizer(object):
    kind = None
    slots = 0
    slot1_init = 0.0

    def hyper(self, lr, t):
        """float[8] for dsb200_optimizer_step: lr, h1, h2, eps, lr_t, grad scale, -, -."""
        raise NotImplementedError


class GradientDescentOptimizer(_Optimizer):
    kind, slots = OPT_SGD, 0

    def __init__(self, learning_rate):
        self.learning_rate = learning_rate

    def hyper(self, lr, t):
        return [lr, 0., 0., 0., lr, 1., 0., 0.]


class MomentumOptimizer(_Optimizer):
    kind, slots = OPT_MOMENTUM, 1
    tf_slot_names = ('Momentum',)                 # names of the slot variables in a tf.train.Saver checkpoint

    def __init__(self, learning_rate, momentum=0.9, use_nesterov=False):
        if use_nesterov:
            raise NotImplementedError('use_nesterov is not used by the reference drivers')
        self.learning_rate, self.momentum = learning_rate, momentum

    def hyper(self, lr, t):
        return [lr, self.momentum, 0., 0., lr, 1., 0., 0.]


class AdamOptimizer(_Optimizer):
    kind, slots = OPT_ADAM, 2
    tf_slot_names = ('Adam', 'Adam_1')            # first / second moment

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.learning_rate, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon

    def hyper(self, lr, t):
        lr_t = lr * (1. - self.beta2 ** t) ** 0.5 / (1. - self.beta1 ** t)
        return [lr, self.beta1, self.beta2, self.epsilon, lr_t, 1., 0., 0.]


class RMSPropOptimizer(_Optimizer):
    kind, slots, slot1_init = OPT_RMSPROP, 2, 1.0
    tf_slot_names = ('RMSProp', 'RMSProp_1')      # mean square, momentum

    def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10):
        self.learning_rate, self.decay, self.momentum, self.epsilon = learning_rate, decay, momentum, epsilon

    def hyper(self, lr, t):
        return [lr, self.decay, self.momentum, self.epsilon, lr, 1., 0., 0.]
