"""Parameter containers for the Point Transformer blocks with the reference's state_dict names
(/root/reference/models/scene_models/pointtransformer.py:9-123).  They hold weights only: execution is in
amb200.scene_engine (fused CUDA kernels); calling them directly is not supported."""
import torch.nn as nn


class PointTransformerLayer(nn.Module):
    def __init__(self, in_planes, out_planes, share_planes=8, nsample=16):
        super().__init__()
        mid = out_planes
        self.mid_planes, self.out_planes, self.share_planes, self.nsample = mid, out_planes, share_planes, nsample
        self.linear_q = nn.Linear(in_planes, mid)
        self.linear_k = nn.Linear(in_planes, mid)
        self.linear_v = nn.Linear(in_planes, out_planes)
        self.linear_p = nn.Sequential(nn.Linear(3, 3), nn.BatchNorm1d(3), nn.ReLU(inplace=True), nn.Linear(3, out_planes))
        self.linear_w = nn.Sequential(nn.BatchNorm1d(mid), nn.ReLU(inplace=True), nn.Linear(mid, mid // share_planes),
                                      nn.BatchNorm1d(mid // share_planes), nn.ReLU(inplace=True),
                                      nn.Linear(out_planes // share_planes, out_planes // share_planes))


class TransitionDown(nn.Module):
    def __init__(self, in_planes, out_planes, stride=1, nsample=16):
        super().__init__()
        self.stride, self.nsample = stride, nsample
        self.linear = nn.Linear((3 if stride != 1 else 0) + in_planes, out_planes, bias=False)
        self.bn = nn.BatchNorm1d(out_planes)


class PointTransformerBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, share_planes=8, nsample=16):
        super().__init__()
        self.linear1 = nn.Linear(in_planes, planes, bias=False)
        self.bn1 = nn.BatchNorm1d(planes)
        self.transformer2 = PointTransformerLayer(planes, planes, share_planes, nsample)
        self.bn2 = nn.BatchNorm1d(planes)
        self.linear3 = nn.Linear(planes, planes * self.expansion, bias=False)
        self.bn3 = nn.BatchNorm1d(planes * self.expansion)


class TransitionUp(nn.Module):
    """pointtransformer.py:72-99 parameters: head form (out_planes None) = linear1(2c->c)+BN, linear2(c->c);
    fusion form = linear1(out->out)+BN, linear2(in->out)+BN."""

    def __init__(self, in_planes, out_planes=None):
        super().__init__()
        self.is_head = out_planes is None
        if out_planes is None:
            self.linear1 = nn.Sequential(nn.Linear(2 * in_planes, in_planes), nn.BatchNorm1d(in_planes), nn.ReLU(inplace=True))
            self.linear2 = nn.Sequential(nn.Linear(in_planes, in_planes), nn.ReLU(inplace=True))
        else:
            self.linear1 = nn.Sequential(nn.Linear(out_planes, out_planes), nn.BatchNorm1d(out_planes), nn.ReLU(inplace=True))
            self.linear2 = nn.Sequential(nn.Linear(in_planes, out_planes), nn.BatchNorm1d(out_planes), nn.ReLU(inplace=True))


class PointTransformerSeg(nn.Module):
    """Frozen scene model of the HUMANISE / novel CDM configs (pointtransformer.py:126-213; cdm.py:436-446,508):
    5-level U-Net, planes [32,64,128,256,512], output [B, N, 32].  Same state_dict names (enc1..5 / dec5..1).
    Executed by amb200.scene_engine.SceneSegEngine (eval-mode BatchNorm folded; the reference freezes it and keeps
    its BatchNorm in eval even under model.train(), utils/training.py:111-116)."""

    def __init__(self, block, blocks, c=6, num_points=8192):
        super().__init__()
        self.num_points, self.c, self.blocks = num_points, c, list(blocks)
        self.in_planes, planes = c, [32, 64, 128, 256, 512]
        share_planes = 8
        stride, nsample = [1, 4, 4, 4, 4], [8, 16, 16, 16, 16]
        for i in range(5):
            setattr(self, f"enc{i + 1}", self._make_enc(block, planes[i], blocks[i], share_planes, stride[i], nsample[i]))
        for i in range(4, -1, -1):
            setattr(self, f"dec{i + 1}", self._make_dec(block, planes[i], 2, share_planes, nsample[i], is_head=(i == 4)))
        self._engine = None

    @property
    def num_groups(self):
        return self.num_points // 256

    def _make_enc(self, block, planes, blocks, share_planes=8, stride=1, nsample=16):
        layers = [TransitionDown(self.in_planes, planes * block.expansion, stride, nsample)]
        self.in_planes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.in_planes, self.in_planes, share_planes, nsample=nsample))
        return nn.Sequential(*layers)

    def _make_dec(self, block, planes, blocks, share_planes=8, nsample=16, is_head=False):
        layers = [TransitionUp(self.in_planes, None if is_head else planes * block.expansion)]
        self.in_planes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.in_planes, self.in_planes, share_planes, nsample=nsample))
        return nn.Sequential(*layers)

    @property
    def engine(self):
        if self._engine is None:
            from amb200.scene_engine import SceneSegEngine
            self._engine = SceneSegEngine(self)
        return self._engine

    def forward(self, pxo):
        """(p [B,N,3], x [B,N,c-3]) -> [B,N,32]  (the form cdm.py:508 uses).  The packed (p, x, o) form is accepted when
        every segment has the same length."""
        import torch
        if len(pxo) == 2:
            p, x = pxo
            return self.engine.forward(p, x)
        if len(pxo) == 3:
            p0, x0, o0 = pxo
            b = o0.numel()
            n = p0.shape[0] // b
            if not torch.equal(o0.cpu().long(), torch.arange(1, b + 1) * n):
                raise ValueError("afford-motion_b200 PointTransformerSeg: packed input needs equal-length segments")
            return self.engine.forward(p0.view(b, n, 3), x0.view(b, n, -1)).reshape(b * n, -1)
        raise ValueError("Input must be (p, x, o) or (p, x)")

    def load_pretrained_weight(self, weight_path: str) -> None:
        """pointtransformer.py:203-213: keep the enc*/dec* entries of the checkpoint."""
        import os
        import torch
        if not os.path.exists(weight_path):
            raise Exception("Can't find pretrained point-transformer weights.")
        model_dict = torch.load(weight_path, map_location="cpu")
        self.load_state_dict({k: v for k, v in model_dict.items() if "enc" in k or "dec" in k})


def pointtransformer_seg_repro(**kwargs) -> PointTransformerSeg:
    return PointTransformerSeg(PointTransformerBlock, [2, 3, 4, 6, 3], **kwargs)
