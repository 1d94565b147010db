"""Flat parameter vector <-> the reference's GymEnvModel state_dict.

The reference saves ``torch.save(elite.state_dict(), logs/<env>/<ts>/saved_models/ep_<n>.pt)``
(learning_strategies/evolution/loop.py:101-104) and test.py:39-40 loads it back into a GymEnvModel;
the keys / shapes / order are those of networks/neural_network.py:12-17.
"""
from collections import OrderedDict

import torch

HID = 32


def param_shapes(obs_dim, act_dim, gru):
    shapes = [("fc1.weight", (HID, obs_dim)), ("fc1.bias", (HID,))]
    if gru:
        shapes += [("gru.weight_ih_l0", (3 * HID, HID)), ("gru.weight_hh_l0", (3 * HID, HID)),
                   ("gru.bias_ih_l0", (3 * HID,)), ("gru.bias_hh_l0", (3 * HID,))]
    shapes += [("fc2.weight", (act_dim, HID)), ("fc2.bias", (act_dim,))]
    return shapes


def flat_to_state_dict(flat, obs_dim, act_dim, gru):
    flat = torch.as_tensor(flat).detach().to("cpu", torch.float32).reshape(-1)
    sd, o = OrderedDict(), 0
    for name, shape in param_shapes(obs_dim, act_dim, gru):
        n = 1
        for s in shape:
            n *= s
        sd[name] = flat[o:o + n].reshape(shape).clone()
        o += n
    if o != flat.numel():
        raise ValueError("flat vector has %d parameters, layout needs %d" % (flat.numel(), o))
    return sd


def state_dict_to_flat(sd, obs_dim, act_dim, gru):
    parts = []
    for name, shape in param_shapes(obs_dim, act_dim, gru):
        t = torch.as_tensor(sd[name]).to(torch.float32)
        if tuple(t.shape) != tuple(shape):
            raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), shape))
        parts.append(t.reshape(-1))
    return torch.cat(parts)
