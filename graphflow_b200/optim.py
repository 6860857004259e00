"""The reference's optimizers on the flat parameter vector, updated on the device through the C-ABI
(`ccn_adam_step`, `ccn_momentum_step`).  State lives in HBM next to the parameters; nothing returns to the host.

Reference: GraphFlow/Adam.h:24-147 (models call `sgd->Learn(learning_rate, nBatch)`, e.g. SMP_beta.h:735,771 -- the
overload whose bias-correction powers advance once per ELEMENT, Adam.h:123,127), GraphFlow/Momentum.h:23-80,
GraphFlow/SGD.h:24-58.  Parameters are registered in the order H, K_l, b_l ..., W (SMP_beta.h:274-280), which is the
order of the flat vector `CCNModelB200` uses."""
import torch


class Adam:
    """Adam::Learn.  `learn(alpha, n_batch)` = `Learn(alpha, nBatch)`; `learn(alpha)` = `Learn(alpha)`."""

    def __init__(self, ctx, params, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.ctx, self.params = ctx, params
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon
        self.m = torch.zeros_like(params)
        self.v = torch.zeros_like(params)
        self.element_updates = 0   # Learn(alpha, nBatch): beta_t advances once per element (Adam.h:123,127)
        self.calls = 0             # Learn(alpha): once per call (Adam.h:82-83)

    def learn(self, grads, alpha, n_batch=None):
        # The reference keeps ONE pair beta1_t / beta2_t for both overloads; mixing them is not supported here.
        if n_batch is None:
            if self.element_updates:
                raise RuntimeError("learn(alpha) after learn(alpha, n_batch): the two overloads share beta_t in the reference")
            self.ctx.adam_step(self.params, grads, self.m, self.v, alpha, 1, self.calls, False, self.beta1, self.beta2,
                               self.epsilon)
            self.calls += 1
        else:
            if self.calls:
                raise RuntimeError("learn(alpha, n_batch) after learn(alpha): the two overloads share beta_t in the reference")
            self.ctx.adam_step(self.params, grads, self.m, self.v, alpha, n_batch, self.element_updates, True, self.beta1,
                               self.beta2, self.epsilon)
            self.element_updates += self.params.numel()
        return self.params


class Momentum:
    """Momentum::Learn (gamma = 0.9 by default); gamma = 0 is SGD::Learn."""

    def __init__(self, ctx, params, gamma=0.9):
        self.ctx, self.params, self.gamma = ctx, params, gamma
        self.moments = torch.zeros_like(params)

    def learn(self, grads, learning_rate, n_batch=1):
        return self.ctx.momentum_step(self.params, grads, self.moments, learning_rate, self.gamma, n_batch)


class SGD(Momentum):
    def __init__(self, ctx, params):
        super().__init__(ctx, params, gamma=0.0)
