"""Masked recurrent state encoder of the policy (pointnav_vo/model_utils/rnns/rnn_state_encoder.py:5-140).

The LSTM/GRU cell itself stays the cuDNN library call (`nn.LSTM`): it is not on the hot path named by
BASELINE.json.  What changes is the sequence path: the reference finds episode boundaries
with a device->host sync per update (`has_zeros ... .cpu()`, :100-111); here the boundary list is computed
from one boolean reduction copied to the host once per call.
"""
import torch
import torch.nn as nn


class RNNStateEncoder(nn.Module):
    def __init__(self, input_size, hidden_size, num_layers=1, rnn_type="GRU"):
        super().__init__()
        self._num_recurrent_layers = num_layers
        self._rnn_type = rnn_type
        self.rnn = getattr(nn, rnn_type)(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers)
        self.layer_init()

    def layer_init(self):
        for name, param in self.rnn.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(param)
            elif "bias" in name:
                nn.init.constant_(param, 0)

    @property
    def num_recurrent_layers(self):
        return self._num_recurrent_layers * (2 if "LSTM" in self._rnn_type else 1)

    def _pack_hidden(self, hidden_states):
        if "LSTM" in self._rnn_type:
            hidden_states = torch.cat([hidden_states[0], hidden_states[1]], dim=0)
        return hidden_states

    def _unpack_hidden(self, hidden_states):
        if "LSTM" in self._rnn_type:
            n = self._num_recurrent_layers
            hidden_states = (hidden_states[0:n], hidden_states[n:])
        return hidden_states

    def _mask_hidden(self, hidden_states, masks):
        if isinstance(hidden_states, tuple):
            return tuple(v * masks for v in hidden_states)
        return masks * hidden_states

    def single_forward(self, x, hidden_states, masks):
        hidden_states = self._unpack_hidden(hidden_states)
        x, hidden_states = self.rnn(x.unsqueeze(0), self._mask_hidden(hidden_states, masks.unsqueeze(0)))
        return x.squeeze(0), self._pack_hidden(hidden_states)

    def seq_forward(self, x, hidden_states, masks):
        n = hidden_states.size(1)
        t = int(x.size(0) / n)
        x = x.view(t, n, x.size(1)).contiguous()
        masks = masks.view(t, n).contiguous()
        # steps (after t = 0) at which any agent starts a new episode
        starts = (masks[1:] == 0.0).any(dim=-1).cpu()
        bounds = [0] + [int(i) + 1 for i in torch.nonzero(starts).flatten().tolist()] + [t]
        hidden_states = self._unpack_hidden(hidden_states)
        outputs = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            scores, hidden_states = self.rnn(x[a:b], self._mask_hidden(hidden_states, masks[a].view(1, -1, 1)))
            outputs.append(scores)
        x = torch.cat(outputs, dim=0).view(t * n, -1).contiguous()
        return x, self._pack_hidden(hidden_states)

    def forward(self, x, hidden_states, masks):
        if x.size(0) == hidden_states.size(1):
            return self.single_forward(x, hidden_states, masks)
        return self.seq_forward(x, hidden_states, masks)
