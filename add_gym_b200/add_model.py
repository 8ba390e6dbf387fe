"""ADDModel: actor, critic and discriminator MLPs, drop-in for the reference ``ADDModel`` / ``PPOModel``
(add/add_model.py:7-50, ppo_model.py:8-67, nets/fc_3layers_1024units.py, nets/fc_2layers_1024units.py,
distribution_gaussian_diag.py:13-58).

* The 22 trainable tensors live in ONE flat fp32 device vector (32-byte aligned segments, reference
  registration order); the ``torch.nn.Linear`` modules only provide the reference's state-dict key names
  (``_actor_layers.{0,2,4}``, ``_action_dist._mean_net``, ``_action_dist._logstd_net``, ``_critic_layers.*``,
  ``_critic_out``, ``_disc_layers.{0,2}``, ``_disc_logits``) and are re-pointed at views of that vector.
  One vector = one AdamW launch and one NCCL all-reduce per optimizer step.
* Initialisation replays the reference's constructor order on the CPU generator (Linear default init,
  zero biases, U(+-0.01) actor head, U(+-1) discriminator head), so the same seed gives the same weights.
* ``eval_actor / eval_critic / eval_disc`` run the library's dense-layer kernels, never torch.nn.functional.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

_NETS = {
    "fc_3layers_1024units": [1024, 1024, 512],
    "fc_2layers_1024units": [1024, 512],
    "fc_2layers_512units": [512, 256],
    "fc_2layers_256units": [256, 128],
    "fc_2layers_128units": [128, 64],
    "fc_2layers_64units": [64, 32],
}


def _build_net(net_name, in_size):
    assert net_name in _NETS, "Unsupported net: {}".format(net_name)
    layers = []
    for out_size in _NETS[net_name]:
        lin = torch.nn.Linear(in_size, out_size)
        torch.nn.init.zeros_(lin.bias)
        layers += [lin, torch.nn.ReLU()]
        in_size = out_size
    return torch.nn.Sequential(*layers), in_size


class DistributionGaussianDiag:
    """Value object returned by eval_actor (distribution_gaussian_diag.py:61-117); plain tensor math."""

    def __init__(self, mean, logstd):
        self._mean, self._logstd = mean, logstd
        self._std = torch.exp(logstd)
        self._dim = mean.shape[-1]

    stddev = property(lambda s: s._std)
    logstd = property(lambda s: s._logstd)
    mean = property(lambda s: s._mean)
    mode = property(lambda s: s._mean)

    def sample(self):
        noise = torch.normal(torch.zeros_like(self._mean), torch.ones_like(self._std))
        return self._mean + self._std * noise

    def log_prob(self, x):
        logp = -0.5 * torch.sum(torch.square((x - self._mean) / self._std), dim=-1)
        logp += -0.5 * self._dim * np.log(2.0 * np.pi) - torch.sum(self._logstd, dim=-1)
        return logp

    def entropy(self):
        return torch.sum(self._logstd, dim=-1) + 0.5 * self._dim * np.log(2.0 * np.pi * np.e)

    def param_reg(self):
        return torch.sum(torch.square(self._mean), dim=-1)


class _ActionDist(torch.nn.Module):
    def __init__(self, in_size, out_size, init_std, init_output_scale):
        super().__init__()
        self._mean_net = torch.nn.Linear(in_size, out_size)
        torch.nn.init.uniform_(self._mean_net.weight, -init_output_scale, init_output_scale)
        torch.nn.init.zeros_(self._mean_net.bias)
        self._logstd_net = torch.nn.Parameter(torch.zeros(out_size, dtype=torch.float32), requires_grad=False)
        torch.nn.init.constant_(self._logstd_net, np.log(init_std))


def _round8(n):
    # segments start on 8-element boundaries: 32 bytes in fp32, 16 bytes in the bf16 shadow (TMA needs 16)
    return (n + 7) & ~7


class ADDModel(torch.nn.Module):
    def __init__(self, config, env, obs_shape, action_space, d_obs_shape, device="cuda"):
        super().__init__()
        self._activation = torch.nn.ReLU
        obs_dim = int(np.prod(obs_shape))
        disc_dim = int(np.prod(d_obs_shape))
        act_dim = int(action_space.shape[0])
        assert config["actor_std_type"] == "FIXED", "the B200 path implements the FIXED action std of add_g1.yaml"
        # reference construction order: actor layers, action dist, critic layers, critic out, disc layers, disc logits
        self._actor_layers, a_out = _build_net(config["actor_net"], obs_dim)
        self._action_dist = _ActionDist(a_out, act_dim, config["action_std"], config["actor_init_output_scale"])
        self._critic_layers, c_out = _build_net(config["critic_net"], obs_dim)
        self._critic_out = torch.nn.Linear(c_out, 1)
        torch.nn.init.zeros_(self._critic_out.bias)
        self._disc_layers, d_out = _build_net(config["disc_net"], disc_dim)
        self._disc_logits = torch.nn.Linear(d_out, 1)
        torch.nn.init.uniform_(self._disc_logits.weight, -1.0, 1.0)
        torch.nn.init.zeros_(self._disc_logits.bias)
        self.obs_dim, self.disc_dim, self.act_dim = obs_dim, disc_dim, act_dim
        prec = config.get("mlp_precision", "fp32")
        self.precision = _lib.PRECISIONS[prec]
        if prec in ("tf32x3", "tf32") and not _lib.has_legacy_kernels():
            raise _lib.AddkError("mlp_precision %r is a superseded mode: build csrc with `make LEGACY=1` and set "
                                 "ADDK_LIB=.../libaddk_legacy.so (default modes: f16x3, bf16, fp32)" % prec)
        self._pack(torch.device(device))

    # ---- flat parameter vector -----------------------------------------------------------------------------
    def trainable(self):
        a, c, d = self._actor_layers, self._critic_layers, self._disc_layers
        lins = [a[0], a[2], a[4], self._action_dist._mean_net, c[0], c[2], c[4], self._critic_out, d[0], d[2],
                self._disc_logits]
        names = ["a_w0", "a_b0", "a_w1", "a_b1", "a_w2", "a_b2", "a_wm", "a_bm", "c_w0", "c_b0", "c_w1", "c_b1",
                 "c_w2", "c_b2", "c_wo", "c_bo", "d_w0", "d_b0", "d_w1", "d_b1", "d_wl", "d_bl"]
        tensors = []
        for l in lins:
            tensors += [l.weight, l.bias]
        return names, tensors

    def _pack(self, device):
        assert len(self._actor_layers) == 6 and len(self._critic_layers) == 6 and len(self._disc_layers) == 4, \
            "the fused update expects 3 hidden layers for actor/critic and 2 for the discriminator"
        names, tensors = self.trainable()
        self.offsets, off = {}, 0
        for n, t in zip(names, tensors):
            self.offsets["o_" + n] = off
            off += _round8(t.numel())
        self.num_params = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=device)
        for n, t in zip(names, tensors):
            o = self.offsets["o_" + n]
            view = self.flat[o:o + t.numel()].view(t.shape)
            view.copy_(t.data)
            t.data = view
            t.grad = self.flat_grad[o:o + t.numel()].view(t.shape)
        self._action_dist._logstd_net.data = self._action_dist._logstd_net.data.to(device)
        self.hidden = ([self._actor_layers[i].out_features for i in (0, 2, 4)],
                       [self._disc_layers[i].out_features for i in (0, 2)])
        assert [self._critic_layers[i].out_features for i in (0, 2, 4)] == self.hidden[0]

    def rebind_storage(self, flat, flat_grad):
        """Move the flat parameter / gradient vectors into caller-provided device memory (the peer-mapped buffers of
        _lib.P2PExchange): values are copied, every parameter and its .grad become views of the new vectors."""
        assert flat.numel() == self.num_params and flat_grad.numel() == self.num_params
        flat.copy_(self.flat)
        flat_grad.copy_(self.flat_grad)
        names, tensors = self.trainable()
        for n, t in zip(names, tensors):
            o = self.offsets["o_" + n]
            t.data = flat[o:o + t.numel()].view(t.shape)
            t.grad = flat_grad[o:o + t.numel()].view(t.shape)
        self.flat, self.flat_grad = flat, flat_grad

    # ---- reference API ---------------------------------------------------------------------------------------
    def _linear(self, x, lin, relu):
        x = x.contiguous()
        y = torch.empty(x.shape[0], lin.out_features, dtype=torch.float32, device=x.device)
        a = _lib.AddkGemmArgs(A=x.data_ptr(), lda=x.shape[1], B=lin.weight.data_ptr(), ldb=lin.in_features,
                              C=y.data_ptr(), ldc=lin.out_features, M=x.shape[0], N=lin.out_features,
                              K=lin.in_features, bias=lin.bias.data_ptr(), a_mean=None, a_std=None,
                              relu_mask_src=None, ld_mask=0, trans_a=0, trans_b=1, relu=int(relu), split_k=1,
                              accumulate=0, slab_stride=0, A16=None, B16=None, C16=None)
        _lib.ptr(x)
        _lib.check(_lib.lib().addk_gemm(_lib.stream(), C.byref(a), C.c_int(self.precision)), "addk_gemm")
        return y

    def _mlp(self, x, layers):
        for i in range(0, len(layers), 2):
            x = self._linear(x, layers[i], relu=True)
        return x

    def eval_actor(self, obs):
        h = self._mlp(obs, self._actor_layers)
        mean = self._linear(h, self._action_dist._mean_net, relu=False)
        logstd = torch.broadcast_to(self._action_dist._logstd_net, mean.shape)
        return DistributionGaussianDiag(mean=mean, logstd=logstd)

    def eval_critic(self, obs):
        return self._linear(self._mlp(obs, self._critic_layers), self._critic_out, relu=False)

    def eval_disc(self, disc_obs):
        return self._linear(self._mlp(disc_obs, self._disc_layers), self._disc_logits, relu=False)

    def get_disc_logit_weights(self):
        return torch.flatten(self._disc_logits.weight)

    def get_disc_weights(self):
        return [torch.flatten(self._disc_layers[0].weight), torch.flatten(self._disc_layers[2].weight),
                torch.flatten(self._disc_logits.weight)]

    def get_actor_params(self):
        return list(self._actor_layers.parameters()) + [self._action_dist._mean_net.weight,
                                                        self._action_dist._mean_net.bias]

    def get_critic_params(self):
        return list(self._critic_layers.parameters()) + list(self._critic_out.parameters())
