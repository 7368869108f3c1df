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
