"""CaptionModel base class (reference: caption_src/CaptionModel.py:18-128).

`beam_search(state, logprobs, feat, pos_feat, opt=...)` keeps the reference's signature and result
(list of `beam_size` dicts {'seq','logps','p'} sorted by score) for callers that drive a single video
by hand.  It advances the beams with the CUDA word step (`get_logprobs_state` -> xg_decode_step); only
the O(beam^2) candidate bookkeeping runs on the host, with the PyTorch-0.3 scalar rules the reference
was written for.  `SAModel.sample_beam` does NOT go through this method: it runs the whole batch on
the device in one call (xg_sample_beam).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


class CaptionModel(nn.Module):
    def __init__(self):
        super(CaptionModel, self).__init__()

    def beam_search(self, state, logprobs, feat, pos_feat, *args, **kwargs):
        opt = kwargs["opt"]
        beam_size = opt.get("beam_size", 5)
        T = self.seq_length
        dev = logprobs.device
        beam_seq = np.zeros((T, beam_size), dtype=np.int64)
        beam_lp = np.zeros((T, beam_size), dtype=np.float32)
        beam_sum = [0.0] * beam_size
        done_beams = []
        feat_ = feat.unsqueeze(0).expand(beam_size, feat.size(0), feat.size(1)).contiguous()
        pos_ = pos_feat.unsqueeze(0).expand(beam_size, pos_feat.size(0)).contiguous()
        for t in range(T):
            lp = logprobs.detach().float().cpu().numpy().copy()
            lp[:, 1] -= 1000.0                                    # CaptionModel.py:94
            cols = min(beam_size, lp.shape[1])
            rows = 1 if t == 0 else beam_size
            # per-row descending order, lowest index first on ties (only the first `cols` are needed)
            order = np.argsort(-lp, axis=1, kind="stable")[:, :cols]
            cand = []
            for c in range(cols):                                 # :45-50, column-major
                for q in range(rows):
                    r = float(lp[q, order[q, c]])
                    cand.append((beam_sum[q] + r, int(order[q, c]), q, r))
            cand.sort(key=lambda x: -x[0])                        # :51, stable
            prev_seq, prev_lp = beam_seq[:t].copy(), beam_lp[:t].copy()
            parents = []
            new_sum = list(beam_sum)
            for vix in range(beam_size):                          # :60-74
                p, c, q, r = cand[vix]
                if t >= 1:
                    beam_seq[:t, vix] = prev_seq[:, q]
                    beam_lp[:t, vix] = prev_lp[:, q]
                parents.append(q)
                beam_seq[t, vix] = c
                beam_lp[t, vix] = r
                new_sum[vix] = float(np.float32(p))
            beam_sum = new_sum
            idx = torch.as_tensor(parents, device=dev)
            state = [tuple(s.index_select(1, idx) for s in state[0]), tuple(s.index_select(1, idx) for s in state[1])]
            for vix in range(beam_size):                          # :108-118
                if beam_seq[t, vix] == 0 or t == T - 1:
                    done_beams.append({"seq": torch.from_numpy(beam_seq[:, vix].copy()),
                                       "logps": torch.from_numpy(beam_lp[:, vix].copy()),
                                       "p": beam_sum[vix]})
                    beam_sum[vix] = -1000.0
            it = torch.from_numpy(beam_seq[t].copy()).to(dev)
            logprobs, state = self.get_logprobs_state(it, feat_, pos_, *(args + (state,)))   # :125
        done_beams = sorted(done_beams, key=lambda x: -x["p"])[:beam_size]                   # :127
        return done_beams
