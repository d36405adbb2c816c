"""Masked recurrent state encoder of the policy (pointnav_vo/model_utils/rnns/rnn_state_encoder.py:5-140).

The LSTM/GRU cell itself stays the cuDNN library call (`nn.LSTM`): it is not on the hot path named by
BASELINE.json.  What changes is the sequence path: the reference finds episode boundaries
with a device->host sync per update (`has_zeros ... .cpu()`, :100-111) and calls the RNN once per piece of the batch;
here every env is cut at its own resets and ONE packed-sequence call runs them all (`seq_forward`).
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
        """[T * N] rows (time-major) through the RNN with per-env state resets where masks == 0.

        The reference cuts the WHOLE batch at every step at which any env resets and calls the RNN once per piece
        (rnn_state_encoder.py:100-138): with 64 envs per minibatch nearly every step is a cut, i.e. ~T calls of one step
        each (~8000 host-launched kernels per PPO update, the update was host-bound).  Here every env's trajectory is cut at
        ITS OWN resets only, the resulting variable-length sequences (N + number of resets) are packed, longest first, and
        the RNN runs ONCE over the PackedSequence; a reset is a fresh sequence with a zero initial state, which is what
        multiplying the carried state by mask = 0 does.  One device->host copy of the mask pattern per call."""
        import numpy as np
        from torch.nn.utils.rnn import PackedSequence

        n = hidden_states.size(1)
        t = int(x.size(0) / n)
        dev = x.device
        x = x.view(t * n, x.size(1))
        masks = masks.view(t, n)
        zero = (masks == 0.0).cpu().numpy()           # the one sync of this call
        starts = zero.copy()
        starts[0, :] = True                           # every env opens a sequence at t = 0
        env, st = np.nonzero(starts.T)                # env-major, time ascending within an env
        last_of_env = np.r_[env[1:] != env[:-1], True]
        end = np.r_[st[1:], t]
        end[last_of_env] = t
        length = end - st
        order = np.argsort(-length, kind="stable")
        env_s, st_s, len_s = env[order], st[order], length[order]
        n_seq, max_len = len(order), int(len_s[0])
        steps = np.arange(max_len)[:, None]
        valid = steps < len_s[None, :]                # row j: the first batch_sizes[j] (longest) sequences
        rows = ((st_s[None, :] + steps) * n + env_s[None, :])[valid]   # packed order -> row of the [T * N] layout
        batch_sizes = torch.from_numpy(valid.sum(1).astype(np.int64))
        rows_t = torch.from_numpy(rows.astype(np.int64)).to(dev, non_blocking=True)
        inv_t = torch.from_numpy(np.argsort(rows, kind="stable").astype(np.int64)).to(dev, non_blocking=True)
        env_t = torch.from_numpy(env_s.astype(np.int64)).to(dev, non_blocking=True)
        # initial state: the carried one (times masks[0], as the reference multiplies) for sequences opening at t = 0, else 0
        carried = torch.from_numpy((st_s == 0).astype(np.float32)).to(dev, non_blocking=True) * masks[0].index_select(0, env_t)
        h = self._unpack_hidden(hidden_states)
        init = lambda v: v.index_select(1, env_t) * carried.view(1, n_seq, 1)  # noqa: E731
        h0 = tuple(init(v) for v in h) if isinstance(h, tuple) else init(h)
        out, hn = self.rnn(PackedSequence(x.index_select(0, rows_t), batch_sizes), h0)
        y = out.data.index_select(0, inv_t)           # back to time-major [T * N] rows
        # final state of env e = final state of its last sequence
        pos = np.empty(n_seq, dtype=np.int64)
        pos[order] = np.arange(n_seq)
        final_t = torch.from_numpy(pos[np.nonzero(last_of_env)[0]]).to(dev, non_blocking=True)   # env order (env is ascending)
        pick = lambda v: v.index_select(1, final_t)   # noqa: E731
        hn = tuple(pick(v) for v in hn) if isinstance(hn, tuple) else pick(hn)
        return y, self._pack_hidden(hn)

    def forward(self, x, hidden_states, masks):
        if x.size(0) == hidden_states.size(1):
            return self.single_forward(x, hidden_states, masks)
        return self.seq_forward(x, hidden_states, masks)
