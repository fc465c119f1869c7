"""oracle/keras_graph.py — run a Keras 2.0.x functional-model JSON (the files under the reference's data/model/) layer
by layer in fp32 PyTorch.  TEST INFRASTRUCTURE.

oracle/model.py restates agent/model.py:32-83 by reading the code.  This module takes the topology from the reference's
own ARTIFACT instead — the `model_*_config.json` that Keras wrote for the shipped networks — and executes exactly the
layers and connections listed there (InputLayer, Conv2D, BatchNormalization, Activation, Add, Flatten, Dense) with the
semantics Keras 2.0.8 documents for them.  tests/test_oracle_vs_reference.py checks that both give the same numbers on
the shipped 192x10 weights (and on the 28-plane config), which pins the restated topology, layer names, epsilon,
activations and flatten order to the reference's files.  The layer arithmetic itself remains a restatement: TensorFlow
is not installable here, so NN parity stays "unpinned by any reference-run vector" (DESIGN.md).
"""
import json

import numpy as np
import torch
import torch.nn.functional as F


def _weight(weights, layer, name):
    return torch.as_tensor(np.asarray(weights[f"{layer}/{name}"]), dtype=torch.float32)


def _activation(x, kind):
    if kind in (None, "linear"):
        return x
    if kind == "relu":
        return F.relu(x)
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "softmax":
        return torch.softmax(x, dim=-1)
    raise NotImplementedError(kind)


def run(config_path, weights, planes):
    """weights: "<layer>/<weight>" -> array (what cczero_b200.keras_h5.read_keras_weights returns).  Returns the model
    outputs in the order of `output_layers`."""
    with open(config_path, "rt") as f:
        cfg = json.load(f)
    cfg = cfg.get("config", cfg)
    values = {}
    x_in = torch.as_tensor(np.asarray(planes), dtype=torch.float32)
    with torch.no_grad():
        for layer in cfg["layers"]:
            kind, name, c = layer["class_name"], layer["name"], layer["config"]
            ins = [values[n[0]] for node in layer["inbound_nodes"] for n in node]
            if kind == "InputLayer":
                assert list(x_in.shape[1:]) == list(c["batch_input_shape"][1:]), (x_in.shape, c["batch_input_shape"])
                out = x_in
            elif kind == "Conv2D":
                assert c["data_format"] == "channels_first" and c["padding"] in ("same", "valid") and tuple(c["strides"]) == (1, 1)
                k = _weight(weights, name, "kernel").permute(3, 2, 0, 1).contiguous()           # HWIO -> OIHW
                assert k.shape[0] == c["filters"] and list(k.shape[2:]) == list(c["kernel_size"])
                out = F.conv2d(ins[0], k, bias=_weight(weights, name, "bias") if c["use_bias"] else None,
                               padding=(c["kernel_size"][0] // 2, c["kernel_size"][1] // 2) if c["padding"] == "same" else 0)
                out = _activation(out, c.get("activation"))
            elif kind == "BatchNormalization":
                assert c["axis"] == 1
                shape = (1, -1, 1, 1)
                mean, var = _weight(weights, name, "moving_mean"), _weight(weights, name, "moving_variance")
                out = (ins[0] - mean.view(shape)) / torch.sqrt(var.view(shape) + c["epsilon"])
                if c.get("scale", True):
                    out = out * _weight(weights, name, "gamma").view(shape)
                if c.get("center", True):
                    out = out + _weight(weights, name, "beta").view(shape)
            elif kind == "Activation":
                out = _activation(ins[0], c["activation"])
            elif kind == "Add":
                out = ins[0]
                for other in ins[1:]:
                    out = out + other
            elif kind == "Flatten":
                out = ins[0].reshape(ins[0].shape[0], -1)
            elif kind == "Dense":
                out = ins[0] @ _weight(weights, name, "kernel")
                if c["use_bias"]:
                    out = out + _weight(weights, name, "bias")
                out = _activation(out, c.get("activation"))
            else:
                raise NotImplementedError(kind)
            values[name] = out
    return [values[o[0]].numpy() for o in cfg["output_layers"]]
