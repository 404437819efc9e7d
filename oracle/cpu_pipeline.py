"""CPU restatement of one frame of the measured hot path, assembled from oracle.py.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (bench.py cpu_baseline and --impl reference,
__graft_entry__.smoke, tests).  This is the "restated reference CPU path" of BASELINE.md 3:
mmcv's single-threaded CPU voxelization loop (C), spconv's native gather-mm-scatter with
torch.mm on all host threads, the RoIAlign C loop and torch CPU linear / bmm / layer_norm.
"""
import numpy as np
import torch

from . import oracle as O

ENC_ARGS = {
    'nusc': dict(in_channels=5, output_channels=128, encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
                 encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock'),
    'waymo': dict(in_channels=5, output_channels=128, encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
                  encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock'),
    'kitti': dict(in_channels=4),
}


def encoder_plan(kind):
    a = dict(base_channels=16, output_channels=128, encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
             encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)), block_type='conv_module')
    a.update(ENC_ARGS[kind])
    return O.encoder_layer_plan(a['in_channels'], a['base_channels'], a['output_channels'], a['encoder_channels'],
                                a['encoder_paddings'], a['block_type'])


def vfe_params(sd):
    p = {}
    for k, v in sd.items():
        if k.startswith('cen2point_pos_enc.'):
            p['pos.' + k[len('cen2point_pos_enc.'):]] = v
        elif k.startswith('vfe_layers.'):
            p['vfe.' + k[len('vfe_layers.'):]] = v
    return p


def encode(state, kind, geom, points):
    if kind == 'nusc':
        v, c, n, _ = O.hard_voxelize(points, geom['voxel_size'], geom['pc_range'], geom['max_points'], geom['max_voxels'])
        feats = O.hard_simple_vfe(v, n, points.shape[1])
        coors = np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)
    else:
        pts, coors_pts = O.detector_voxelize_dynamic([points], geom['voxel_size'], geom['pc_range'])
        feats, coors = O.dynamic_vfe_custom(vfe_params(state['vfe']), pts, coors_pts, geom['voxel_size'], geom['pc_range'])
    return O.sparse_encoder(state['encoder'], encoder_plan(kind), feats, coors, 1, geom['sparse_shape'])


def region_stages(state, geom, d, strides=(8, 16, 32, 64), istrides=(4, 8, 16, 32)):
    prop = state['prop0']
    for s, boxes in enumerate(state['stage_boxes']):
        b = boxes.copy()
        if state['fuse'] is not None:
            img = O.img_roi_feats(state['img_feats'], b, state['lidar2img'][None], geom['pc_range'], list(istrides))
            pts = O.points_roi_feats(state['bev_feats'], b, geom['pc_range'], geom['voxel_size'], list(strides))
            roi = O.fusion_proj(img, pts, state['fuse'][s]['weight'], state['fuse'][s]['bias'])
        else:
            roi = O.points_roi_feats(state['bev_feats'], b, geom['pc_range'], geom['voxel_size'], list(strides))
        prop = O.dynamic_conv(state['dynconv'][s], prop, roi, d)
    return prop


def full_chain(state, bev):
    """SECONDCustom -> FPN -> SRFDetHead (DPG, chained stages) -> decode on the dense BEV map (scope 'full')."""
    cfg = state['head_cfg']
    feats = O.second_custom(state['backbone'], bev, [5, 5], [1, 2])
    pyramid = O.fpn(state['neck'], feats, 4, extra_convs=cfg['extra_convs'])
    img = state.get('img_feats')
    l2i = state['lidar2img'][None] if img is not None else None
    trace = {}
    logits, boxes = O.srfdet_head_forward(state['head'], img, pyramid, l2i, cfg, trace=trace)
    scores, dec = O.decode_boxes(logits[-1], boxes[-1])
    return np.concatenate([dec[0], scores[0]], axis=1), dict(pyramid=pyramid, logits=logits, boxes=boxes, trace=trace)


def run_frame(state, kind, geom, d, points):
    torch.set_num_threads(max(1, torch.get_num_threads()))
    bev = encode(state, kind, geom, points)
    if state.get('scope') == 'full':
        return bev, full_chain(state, bev)[0]
    return bev, region_stages(state, geom, d)
