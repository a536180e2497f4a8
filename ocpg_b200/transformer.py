"""OCPG's DeformableTransformer, re-hosted end to end on the B200 path: a drop-in for the class of the same name in
models/deformable_transformer.py (:26-217) -- same constructor arguments, same parameter / sub-module names
(``encoder``, ``decoder``, ``level_embed``, ``reference_points``), same ``forward(srcs, tgt, masks, pos_embeds, query_embed)``
and the same 7-tuple result -- built from the pieces of SURVEY.md section 8f:

    flatten_levels (:149-169)  ->  DeformableTransformerEncoder (:220-290)  ->  reference points of the queries (:188-195)
    ->  DeformableTransformerDecoder (:293-398)  ->  unflatten_levels (:205-212)

Only the configuration OCPG runs is covered: ``two_stage=False`` (opts.py:30: "NOTE: must be false"); asking for the
two-stage variant raises.  Nothing here synchronises with the host except, once per distinct set of shapes, the checks the
reference makes on every call.
"""
from __future__ import annotations

import torch
from torch import nn

from .decoder import DeformableTransformerDecoder, DeformableTransformerDecoderLayer
from .encoder import DeformableTransformerEncoder, DeformableTransformerEncoderLayer
from .flatten import flatten_levels, unflatten_levels
from .modules import MSDeformAttn


class DeformableTransformer(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=1024, dropout=0.1,
                 activation="relu", return_intermediate_dec=False, num_feature_levels=4, dec_n_points=4, enc_n_points=4,
                 two_stage=False, two_stage_num_proposals=300, fused=True):
        super().__init__()
        if two_stage:
            raise NotImplementedError("two_stage=True is not part of OCPG's configuration (opts.py:30) and is not re-hosted")
        self.d_model, self.nhead, self.dropout = d_model, nhead, dropout
        self.two_stage, self.two_stage_num_proposals = two_stage, two_stage_num_proposals
        self.num_feature_level = num_feature_levels
        encoder_layer = DeformableTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                                          nhead, enc_n_points, fused=fused)
        self.encoder = DeformableTransformerEncoder(encoder_layer, num_encoder_layers)
        decoder_layer = DeformableTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels,
                                                          nhead, dec_n_points, fused=fused)
        self.decoder = DeformableTransformerDecoder(decoder_layer, num_decoder_layers, return_intermediate_dec)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self.reference_points = nn.Linear(d_model, 2)
        self._reset_parameters()

    def _reset_parameters(self):                                                  # :65-75
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.xavier_uniform_(self.reference_points.weight.data, gain=1.0)
        nn.init.constant_(self.reference_points.bias.data, 0.0)
        nn.init.normal_(self.level_embed)

    @staticmethod
    def get_valid_ratio(mask):                                                    # :125-133
        _, H, W = mask.shape
        valid_H = torch.sum(~mask[:, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H], -1)

    def forward(self, srcs, tgt, masks, pos_embeds, query_embed=None):
        """srcs / pos_embeds: per level (b*t, c, h_l, w_l); masks: per level (b*t, h_l, w_l) bool, True = padding;
        tgt (b, t, q, c); query_embed (q, c).  Returns (hs, memory_features, init_reference_out, inter_references_out,
        None, None, inter_samples) like the reference (:217)."""
        assert query_embed is not None                                             # :136 with two_stage False
        src_flatten, lvl_pos_embed_flatten, spatial_shapes, level_start_index = flatten_levels(
            list(srcs), list(pos_embeds), self.level_embed)                        # :149-169
        mask_flatten = torch.cat([m.flatten(1) for m in masks], 1)
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)   # :170
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_embed_flatten,
                              mask_flatten)                                        # :173
        b, t, q, c = tgt.shape                                                     # :190-195
        tgt = tgt.reshape(b * t, q, c)
        query_embed = query_embed[None, None].expand(b, t, -1, -1).flatten(0, 1)
        reference_points = self.reference_points(query_embed).sigmoid()
        init_reference_out = reference_points
        hs, inter_references, inter_samples = self.decoder(tgt, reference_points, memory, spatial_shapes, level_start_index,
                                                           valid_ratios, query_embed, mask_flatten)                    # :200
        shapes = [tuple(s.shape[2:]) for s in srcs[:self.num_feature_level - 1]]
        memory_features = unflatten_levels(memory, shapes)                         # :205-212
        return hs, memory_features, init_reference_out, inter_references, None, None, inter_samples


def build_deforamble_transformer(args):
    """The reference's factory, name and all (:409-423)."""
    return DeformableTransformer(
        d_model=args.hidden_dim, nhead=args.nheads, num_encoder_layers=args.enc_layers, num_decoder_layers=args.dec_layers,
        dim_feedforward=args.dim_feedforward, dropout=args.dropout, activation="relu", return_intermediate_dec=True,
        num_feature_levels=args.num_feature_levels, dec_n_points=args.dec_n_points, enc_n_points=args.enc_n_points,
        two_stage=args.two_stage, two_stage_num_proposals=args.num_queries)
