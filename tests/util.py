import numpy as np
import torch


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'max rel err' of BASELINE.json's north_star)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def cuda(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def randomize_bn_(module, seed):
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def nbr_to_pairs(nbr, n_out):
    """(kvol, cap) neighbour table -> list of (in_rows, out_rows) per offset (sorted by out row)."""
    out = []
    for k in range(nbr.shape[0]):
        col = nbr[k, :n_out]
        o = np.nonzero(col >= 0)[0].astype(np.int32)
        out.append((col[o].astype(np.int32), o))
    return out


def box_errors(last, ref, pc):
    """Errors of the chained head outputs: FPN level 0, logits of all stages, boxes of all stages split into
    centres (as a fraction of the range), log sizes and (sin, cos, v) -- absolute differences."""
    span = np.array([pc[3] - pc[0], pc[4] - pc[1], pc[5] - pc[2]], np.float32)
    gb, rb = last['boxes'].cpu().numpy(), ref['boxes']
    stages = last['logits'].shape[0]
    lg, rl = last['logits'].cpu().numpy(), ref['logits']
    return dict(logits_per_stage=[rel_err(lg[i], rl[i]) for i in range(stages)], fpn0=rel_err(last['pyramid'][0].cpu().numpy(), ref['pyramid'][0]), fpn3=rel_err(last['pyramid'][3].cpu().numpy(), ref['pyramid'][3]),
                logits=rel_err(last['logits'].cpu().numpy(), ref['logits']),
                centre_frac=float((np.abs(gb[..., :3] - rb[..., :3]) / span).max()), logsize_abs=float(np.abs(gb[..., 3:6] - rb[..., 3:6]).max()),
                rest_abs=float(np.abs(gb[..., 6:] - rb[..., 6:]).max()))


def teacher_forced(pipe, ref, prec):
    """Every stage of the head fed with the ORACLE's inputs of that stage (FPN pyramid, boxes, proposal features):
    isolates each stage's own error from the amplification of the chained loop."""
    tr = ref['trace']
    head = pipe.head
    pyr = [torch.as_tensor(p).cuda().contiguous(memory_format=torch.channels_last) for p in ref['pyramid']]
    img = pipe.img_feats if pipe.fusion else None
    b0, f0 = head._get_init_proposals(img, pyr, sigmoid_centres=True)
    out = dict(tf_init_boxes=float(np.abs(b0.cpu().numpy() - tr['init_boxes']).max()), tf_init_prop=rel_err(f0[0].cpu().numpy(), tr['init_prop']),
               tf_logits=[], tf_boxes=[], tf_obj=[])
    for s, stage in enumerate(head.head_series_lidar):
        boxes = torch.as_tensor(tr['stage_in'][s][0]).cuda().contiguous()
        prop = torch.as_tensor(tr['stage_in'][s][1]).cuda().contiguous()
        if pipe.fusion:
            lg, pred, obj = stage(img, pyr, boxes, prop, head.roi_extractor_lidar, None, pooler_img=head.roi_extractor_img, precision=prec,
                                  lidar2img=pipe.lidar2img)
        else:
            lg, pred, obj = stage(pyr, boxes, prop, head.roi_extractor_lidar, None, precision=prec)
        rl, rp, ro = tr['stage_out'][s]
        out['tf_logits'].append(rel_err(lg[0].cpu().numpy(), rl))
        out['tf_boxes'].append(float(np.abs(pred[0].cpu().numpy() - rp).max()))
        out['tf_obj'].append(rel_err(obj.cpu().numpy(), ro))
    return out


