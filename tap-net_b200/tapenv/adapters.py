"""Adapters between the reference's network modules and tapenv.DecodeLoop.

`tapenv.install()` keeps model.py byte-identical but inherits its host round trips (model.py:344 `.any()`, :407-412
`.cpu().numpy()`, B Python `add_new_block` calls, :456 `torch.FloatTensor(...).cuda()`).  `drl_actor_step` instead lifts
the NETWORK half of one decode step out of an existing `model.DRL` instance -- its own encoders and pointer, its own
weights, nothing re-implemented -- into the callback `DecodeLoop` drives, so the whole episode stays on the device:

    actor = model.DRL(...); actor.load_state_dict(torch.load(".../actor.pt"))
    env   = tapenv.BatchedContainers([W, H], n, reward_type, heightmap_type, batch_size=B)
    loop  = tapenv.DecodeLoop(env, tapenv.adapters.drl_actor_step(actor), greedy=not actor.training, use_graph=True)
    tour_idx, tour_logp, reward = loop.run(static, dynamic)        # == actor(static, dynamic, [dec_static, dec_dyn]) minus the syncs

Only duck-typed attributes of the module are used (`static_encoder`, `dynamic_encoder`, `static_decoder`, `dynamic_decoder`
/ `decoder`, `pointer`, `decoder_input_type`, `input_type`, `block_dim`), so nothing of the reference is imported here.
"""
import torch


class _State(object):
    """Carried between decode steps.  `last_ptr` is filled in by DecodeLoop after every step (the pointer it just
    applied), which tells the incremental encoder update which rows update_dynamic zeroed."""
    __slots__ = ("static_hidden", "dynamic_hidden", "last_hh", "prev_dynamic", "last_ptr")

    def __init__(self):
        self.static_hidden = self.dynamic_hidden = self.last_hh = self.prev_dynamic = self.last_ptr = None


def _static_part(actor, static):
    it = actor.input_type
    if it == "mul":
        return static[:, 1:-1, :]                      # model.py:318-319, :389-390
    if it == "rot-old":
        return static
    return static[:, 1:, :]                            # model.py:337, :394


def incremental_dynamic_hidden(encoder, hidden_prev, dynamic_prev, dynamic_new, ptr_rows=None):
    """dynamic_encoder(dynamic_new) from dynamic_encoder(dynamic_prev) (model.py:380) when update_dynamic only ZEROED rows
    (pack.py:370-374): the 1x1 Conv1d is linear per column, so

        hidden_new[b,:,j] = hidden_prev[b,:,j] - sum_r W[:, r] * (dynamic_prev - dynamic_new)[b, r, j]

    and the difference tensor is non-zero in at most `update_time` (3) rows per environment -- a rank-3 correction instead of
    the full [hidden x 3n] product.  `ptr_rows` (int64 [B,k]): the rows that were zeroed (real + n*band, pack.py:347-374);
    without it the rows are found from the difference itself (a full read of both tensors -- for checking only).  Result
    equals the full recomputation up to fp32 rounding of a different summation order (tests/test_gpu_model_in_loop.py
    states the tolerance)."""
    W = encoder.conv.weight.squeeze(-1)                # [hidden, rows]
    S = dynamic_prev.shape[2]
    if ptr_rows is None:
        diff = dynamic_prev - dynamic_new              # [B, rows, S], non-zero only in the zeroed rows
        ptr_rows = torch.topk(diff.abs().sum(2), min(3, diff.shape[1]), dim=1).indices     # unchanged rows contribute zeros
        d = torch.gather(diff, 1, ptr_rows.unsqueeze(2).expand(-1, -1, S))
    else:                                              # the zeroed rows of the OLD tensor are the whole difference
        d = torch.gather(dynamic_prev, 1, ptr_rows.unsqueeze(2).expand(-1, -1, S))          # [B,k,S]
    Wk = W.t()[ptr_rows]                               # [B,k,hidden]
    return torch.baddbmm(hidden_prev, Wk.transpose(1, 2), d, alpha=-1.0)


def drl_actor_step(actor, incremental=False):
    """The network half of one iteration of model.DRL.forward's loop (model.py:347-358) as a DecodeLoop callback.

    actor: a `model.DRL` (or rolling.DRL) instance -- its mode (train / eval) is respected (dropout in the pointer).
    incremental=True: `dynamic_hidden` is updated with the rank-3 correction instead of re-running the encoder."""
    dit = actor.decoder_input_type

    def actor_step(static, dynamic, dec_static, dec_dyn, state):
        B = static.shape[0]
        if state is None:
            state = _State()
            state.static_hidden = actor.static_encoder(_static_part(actor, static))           # model.py:318-337
            state.dynamic_hidden = actor.dynamic_encoder(dynamic)                             # model.py:316
        elif incremental and state.prev_dynamic is not None and state.last_ptr is not None:
            rows_total = state.prev_dynamic.shape[1]
            n = rows_total // 3 if rows_total % 3 == 0 and actor.input_type not in ("simple", "rot") else rows_total
            real = torch.gather(static[:, 0, :], 1, state.last_ptr.view(-1, 1)).long()        # pack.py:347
            rows = real + n * torch.arange(rows_total // n, device=real.device).view(1, -1)   # pack.py:370-374
            state.dynamic_hidden = incremental_dynamic_hidden(actor.dynamic_encoder, state.dynamic_hidden,
                                                              state.prev_dynamic, dynamic, rows)
        else:
            state.dynamic_hidden = actor.dynamic_encoder(dynamic)                             # model.py:380
        state.prev_dynamic = dynamic
        ds = dec_static.reshape(B, -1, 1)                                                     # [B,dim,1] (pack.py:228-266)
        dd = dec_dyn.unsqueeze(2) if dec_dyn.dim() == 2 else dec_dyn                          # 2D: [B,enc,1] (model.py:456)
        if dd.dim() == 3 and actor.block_dim == 3:                                            # 3D 'full'/'zero': [B,1,W,L] (model.py:464-465)
            dd = dd.unsqueeze(1)
        if dit == "shape_only":
            decoder_hidden = actor.decoder(ds)                                                # model.py:347-348
        elif dit == "heightmap_only":
            decoder_hidden = actor.dynamic_decoder(dd)
        else:                                                                                 # 'shape_heightmap', model.py:351-354
            decoder_hidden = torch.cat((actor.static_decoder(ds), actor.dynamic_decoder(dd)), 1)
        logits, state.last_hh = actor.pointer(state.static_hidden, state.dynamic_hidden, decoder_hidden, state.last_hh)
        return logits, state

    return actor_step
