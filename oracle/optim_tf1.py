"""Oracle restatement of the TF 1.10 optimizers / schedules the reference's experiment
drivers pass to the model (experiments/cosmo/experiment_deepsphere.py:97-98,
experiments/shrec17/hyperparameters.py:146-150).  TEST INFRASTRUCTURE ONLY.

TensorFlow 1.10.0 (environment.yml:162) is a third-party dependency absent from
/root/reference; the published update rules of tf.train.* are restated.  Parity unpinned.
"""
import torch


def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
    """tf.train.exponential_decay: lr * rate^(step/decay_steps) (continuous unless staircase)."""
    e = global_step / decay_steps
    if staircase:
        e = float(int(e))
    return learning_rate * decay_rate ** e


class Adam:
    """tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps)
    (epsilon is NOT bias-corrected, unlike torch.optim.Adam)."""

    def __init__(self, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.b1, self.b2, self.eps, self.t, self.m, self.v = beta1, beta2, epsilon, 0, {}, {}

    def step(self, params, grads, lr):
        self.t += 1
        lr_t = lr * (1 - self.b2 ** self.t) ** 0.5 / (1 - self.b1 ** self.t)
        for k, p in params.items():
            g = grads[k]
            m = self.m.get(k, torch.zeros_like(p))
            v = self.v.get(k, torch.zeros_like(p))
            m = self.b1 * m + (1 - self.b1) * g
            v = self.b2 * v + (1 - self.b2) * g * g
            self.m[k], self.v[k] = m, v
            params[k] = p - lr_t * m / (v.sqrt() + self.eps)
        return params


class Momentum:
    """tf.train.MomentumOptimizer (use_nesterov=False): a = mom*a + g; p -= lr*a."""

    def __init__(self, momentum=0.9):
        self.mom, self.a = momentum, {}

    def step(self, params, grads, lr):
        for k, p in params.items():
            a = self.mom * self.a.get(k, torch.zeros_like(p)) + grads[k]
            self.a[k] = a
            params[k] = p - lr * a
        return params


class SGD:
    """tf.train.GradientDescentOptimizer."""

    def step(self, params, grads, lr):
        for k, p in params.items():
            params[k] = p - lr * grads[k]
        return params


class RMSProp:
    """tf.train.RMSPropOptimizer: ms (init 1) = d*ms+(1-d)g^2; mom = m*mom + lr*g/sqrt(ms+eps); p -= mom."""

    def __init__(self, decay=0.9, momentum=0.0, epsilon=1e-10):
        self.d, self.m, self.eps, self.ms, self.mom = decay, momentum, epsilon, {}, {}

    def step(self, params, grads, lr):
        for k, p in params.items():
            g = grads[k]
            ms = self.d * self.ms.get(k, torch.ones_like(p)) + (1 - self.d) * g * g
            mom = self.m * self.mom.get(k, torch.zeros_like(p)) + lr * g / (ms + self.eps).sqrt()
            self.ms[k], self.mom[k] = ms, mom
            params[k] = p - mom
        return params
