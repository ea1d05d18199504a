"""Drop-in for the reference's model_convlstm.py hot path: ConvLSTMCell / ConvLSTM (model_convlstm.py:73-236)
and ConvTWACell / ConvTWA (:238-401).  Same constructors, forward signatures, return structures and
``cell_list.{i}.rnn_conv.{weight,bias}`` keys; the recurrence runs in the sm_100a sequence kernels
(uavsal_twa_sequence / uavsal_convlstm_sequence) with the gate math fused into the conv epilogue.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._kernel_module import KernelModule, require_cuda
from .blocks import init_weights

__all__ = ["ConvLSTMCell", "ConvLSTM", "ConvTWACell", "ConvTWA"]


def _check_kernel_size(kernel_size):
    ok = isinstance(kernel_size, tuple) or (isinstance(kernel_size, list) and all(isinstance(e, tuple) for e in kernel_size))
    if not ok:
        raise ValueError("`kernel_size` must be tuple or list of tuples")      # model_convlstm.py:227-230


def _per_layer(param, num_layers):
    return param if isinstance(param, list) else [param] * num_layers


class _CellBase(KernelModule):
    GATES = 1

    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, bias):
        super().__init__()
        self.height, self.width = input_size
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.kernel_size = kernel_size
        self.padding = kernel_size[0] // 2, kernel_size[1] // 2
        self.bias = bias
        self.rnn_conv = nn.Conv2d(input_dim + hidden_dim, self.GATES * hidden_dim, kernel_size, padding=self.padding, bias=bias)

    def _check_supported(self):
        if tuple(self.kernel_size) != (3, 3):
            raise NotImplementedError("only 3x3 recurrent kernels have an sm_100a kernel (the reference path uses (3,3))")
        if self.input_dim % 8 or self.hidden_dim % 8:
            raise NotImplementedError("channel counts must be multiples of 8")

    def wspec(self):
        from .engine import W
        return W(self.rnn_conv.weight, bias=self.rnn_conv.bias, owner=self.rnn_conv)

    def _zeros(self, batch, device):
        return torch.zeros(batch, self.hidden_dim, self.height, self.width, device=device)


class ConvTWACell(_CellBase):
    """h' = i*x + (1-i)*h with i = sigmoid(conv3x3([x, h]))  (model_convlstm.py:276-292)."""
    GATES = 1

    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, bias):
        super().__init__(input_size, input_dim, hidden_dim, kernel_size, bias)
        init_weights(self.rnn_conv, "kaiming_normal", mode="fan_out")            # :274
        if input_dim != hidden_dim:
            raise ValueError("ConvTWA blends x and h element-wise: input_dim must equal hidden_dim")

    def init_hidden(self, batch_size):
        return self._zeros(batch_size, self.rnn_conv.weight.device)             # reference: .cuda() (:294-295)

    def forward(self, input_tensor, cur_state):
        y, h = _run_twa(self, input_tensor.unsqueeze(1), cur_state[0])
        return h


class ConvLSTMCell(_CellBase):
    """i,f,o,g = split(conv3x3([x,h])); c' = s(f)c + s(i)tanh(g); h' = s(o)tanh(c')  (model_convlstm.py:111-126)."""
    GATES = 4

    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, bias):
        super().__init__(input_size, input_dim, hidden_dim, kernel_size, bias)
        init_weights(self.rnn_conv, "xavier_uniform")                            # :109

    def init_hidden(self, batch_size):
        dev = self.rnn_conv.weight.device
        return (self._zeros(batch_size, dev), self._zeros(batch_size, dev))      # reference: .cuda() (:128-130)

    def forward(self, input_tensor, cur_state):
        h, c = cur_state
        y, (h, c) = _run_lstm(self, input_tensor.unsqueeze(1), h, c)
        return h, c


def _run_twa(cell: ConvTWACell, x5: torch.Tensor, h0: torch.Tensor):
    """x5 (b,t,c,h,w), h0 (b,c,h,w) -> (y (b,t,c,h,w), h_last (b,c,h,w)).  One sequence launch per batch element
    (the production path has b = 1, model.py:368)."""
    require_cuda(x5, "ConvTWA")
    cell._check_supported()
    if cell.rnn_conv.bias is not None:
        raise NotImplementedError("ConvTWA with bias=True is not on the UAVSal path (model.py:328-329 uses bias=False)")
    b, t, c, h, w = x5.shape

    def build(plan):
        xin = plan.tensor((b * t, c, h, w))
        hin = plan.tensor((b, c, h, w))
        xb = plan.alloc(b * t * h * w, c)
        hb = plan.alloc(b * h * w, c)
        seq = plan.alloc(b * t * h * w, c)
        plan.pack_nchw(xin, b * t, c, h, w, xb)
        plan.pack_nchw(hin, b, c, h, w, hb)
        emit_twa(plan, cell, xb, hb, seq, b, t, h, w)
        yout = plan.tensor((b * t, c, h, w))
        plan.unpack_nchw(seq, b * t, c, h, w, yout)
        plan.named.update(x_in=xin, h_in=hin, y_out=yout)

    plan = cell._cached_plan((x5.device, "twa", b, t, c, h, w), build)
    plan.named["x_in"].copy_(x5.reshape(b * t, c, h, w))
    plan.named["h_in"].copy_(h0)
    plan.launch()
    y = plan.named["y_out"].view(b, t, c, h, w).clone()
    return y, y[:, -1].clone()


def emit_twa(plan, cell, xb, hb, seq, b, t, h, w):
    """b independent sequences of t frames (sequence-major rows) through one batched sequence op."""
    plan.twa(xb, hb, t, h, w, cell.hidden_dim, cell.wspec(), seq, tag="rnn", batch=b)


def _run_lstm(cell: ConvLSTMCell, x5, h0, c0):
    require_cuda(x5, "ConvLSTM")
    cell._check_supported()
    b, t, cin, h, w = x5.shape
    ch = cell.hidden_dim

    def build(plan):
        xin = plan.tensor((b * t, cin, h, w))
        hin = plan.tensor((b, ch, h, w))
        cin_t = plan.tensor((b, ch, h, w))
        cst = plan.tensor((b, h * w, ch))                     # NHWC fp32 cell state
        xb = plan.alloc(b * t * h * w, cin)
        hb = plan.alloc(b * h * w, ch)
        seq = plan.alloc(b * t * h * w, ch)
        plan.pack_nchw(xin, b * t, cin, h, w, xb)
        plan.pack_nchw(hin, b, ch, h, w, hb)
        plan.lstm(xb, hb, cst, b, t, h, w, cin, ch, cell.wspec(), None, seq, tag="lstm")
        yout = plan.tensor((b * t, ch, h, w))
        plan.unpack_nchw(seq, b * t, ch, h, w, yout)
        plan.named.update(x_in=xin, h_in=hin, c_in=cin_t, c_state=cst, y_out=yout)

    plan = cell._cached_plan((x5.device, "lstm", b, t, cin, h, w), build)
    plan.named["x_in"].copy_(x5.reshape(b * t, cin, h, w))
    plan.named["h_in"].copy_(h0)
    plan.named["c_state"].copy_(c0.permute(0, 2, 3, 1).reshape(b, h * w, ch))   # boundary layout change (NCHW -> NHWC)
    plan.launch()
    y = plan.named["y_out"].view(b, t, ch, h, w).clone()
    c = plan.named["c_state"].view(b, h, w, ch).permute(0, 3, 1, 2).contiguous()
    return y, (y[:, -1].clone(), c)


class _SeqBase(nn.Module):
    CELL = None

    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, num_layers, batch_first=False, bias=True,
                 return_all_layers=False):
        super().__init__()
        _check_kernel_size(kernel_size)
        kernel_size = _per_layer(kernel_size, num_layers)
        hidden_dim = _per_layer(hidden_dim, num_layers)
        if not len(kernel_size) == len(hidden_dim) == num_layers:
            raise ValueError("Inconsistent list length.")
        self.height, self.width = input_size
        self.input_dim, self.hidden_dim, self.kernel_size = input_dim, hidden_dim, kernel_size
        self.num_layers, self.batch_first, self.bias = num_layers, batch_first, bias
        self.return_all_layers = return_all_layers
        self.cell_list = nn.ModuleList(
            self.CELL(input_size=(self.height, self.width), input_dim=input_dim if i == 0 else hidden_dim[i - 1],
                      hidden_dim=hidden_dim[i], kernel_size=kernel_size[i], bias=bias) for i in range(num_layers))

    def _init_hidden(self, batch_size):
        return [cell.init_hidden(batch_size) for cell in self.cell_list]

    _check_kernel_size_consistency = staticmethod(_check_kernel_size)
    _extend_for_multilayer = staticmethod(_per_layer)

    def set_mode(self, precision=None, engine=None):
        for cell in self.cell_list:
            cell.set_mode(precision, engine)
        return self


class ConvTWA(_SeqBase):
    """model_convlstm.py:297-401.  forward(input (t,b,c,h,w) | (b,t,c,h,w), hidden_state=[h per layer]) ->
    (layer_output, [h]) for return_all_layers=False."""
    CELL = ConvTWACell

    def forward(self, input_tensor, hidden_state=None):
        if not self.batch_first:
            input_tensor = input_tensor.permute(1, 0, 2, 3, 4)
        if hidden_state is None:
            hidden_state = self._init_hidden(input_tensor.size(0))
        outs, states = [], []
        cur = input_tensor.contiguous()
        for i, cell in enumerate(self.cell_list):
            cur, h = _run_twa(cell, cur, hidden_state[i])
            outs.append(cur)
            states.append([h])
        if not self.return_all_layers:
            return outs[-1], states[-1]
        return outs, states


class ConvLSTM(_SeqBase):
    """model_convlstm.py:132-236.  hidden_state = [[h, c] per layer]; returns (layer_output, [h, c])."""
    CELL = ConvLSTMCell

    def forward(self, input_tensor, hidden_state=None):
        if not self.batch_first:
            input_tensor = input_tensor.permute(1, 0, 2, 3, 4)
        if hidden_state is None:
            hidden_state = self._init_hidden(input_tensor.size(0))
        outs, states = [], []
        cur = input_tensor.contiguous()
        for i, cell in enumerate(self.cell_list):
            h, c = hidden_state[i]
            cur, (h, c) = _run_lstm(cell, cur, h, c)
            outs.append(cur)
            states.append([h, c])
        if not self.return_all_layers:
            return outs[-1], states[-1]
        return outs, states
