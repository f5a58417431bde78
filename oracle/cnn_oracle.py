"""CPU oracle of the learned motion-cost path (TEST INFRASTRUCTURE ONLY), in fp32 PyTorch functional ops.

Restates, citing the reference (paths under art_planner_motion_cost/):
  * network.CNNpart            src/art_planner_motion_cost/predictor/network_light.py:78-110
  * network.FCpart             src/art_planner_motion_cost/predictor/network_light.py:113-165
  * CostQuery.setMapParams / __call__   src/art_planner_motion_cost/predictor/cost_query.py:27-69
  * the server's map orientation and query centring   scripts/cost_query_server.py:74,160-161
The reference runs the module in fp16 (`predictor.py:22`); parity is judged against the fp32 evaluation of the same
module, which is what this oracle computes. oracle/make_golden_cnn.py checks this restatement
against the reference's own module (imported from /root/reference, fp32, CPU) and commits the golden costs.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

MAP_CLIP = 24          # network_light.py:16
DOWNSAMPLE = 2         # network_light.py:15
BN_EPS = 1e-5


def cnn_input_from_layer(layer: np.ndarray) -> np.ndarray:
    """grid_map layer (rows x cols, index (i,j) at -x,-y) -> network input E[r][c] = layer(rows-1-r, cols-1-c):
    np.rot90(msg.reshape(row, col), 2).transpose() of the column-major message (cost_query_server.py:74)."""
    return np.ascontiguousarray(layer[::-1, ::-1], dtype=np.float32)


class CostNetOracle:
    def __init__(self, state_dict: dict):
        self.p = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in state_dict.items()
                  if not k.endswith("num_batches_tracked")}

    def _conv_bn(self, x, conv, bn):
        p = self.p
        y = F.conv2d(x, p[conv + ".weight"])
        return F.batch_norm(y, p[bn + ".running_mean"], p[bn + ".running_var"], p[bn + ".weight"], p[bn + ".bias"],
                            training=False, eps=BN_EPS)

    @torch.no_grad()
    def features(self, E: np.ndarray) -> torch.Tensor:
        """network.CNNpart (network_light.py:78-110); E [rows, cols] -> [48, (rows-48)/2, (cols-48)/2]."""
        t = torch.as_tensor(E, dtype=torch.float32)[None, None]
        t = self._conv_bn(t, "init_conv1", "init_conv1_bn")
        t = F.leaky_relu(self._conv_bn(t, "init_conv2", "init_conv2_bn"), 0.3)
        t = F.max_pool2d(t, (2, 2), stride=2)
        t = F.leaky_relu(self._conv_bn(t, "init_conv3", "init_conv3_bn"), 0.3)
        t = F.leaky_relu(self._conv_bn(t, "init_conv4", "init_conv4_bn"), 0.3)
        t = F.max_pool2d(t, (3, 3), stride=1)
        t = F.leaky_relu(self._conv_bn(t, "init_conv5", "init_conv5_bn"), 0.3)
        t = F.leaky_relu(self._conv_bn(t, "init_flatten", "init_flatten_bn"), 0.3)
        return t[0]   # dropout is the identity in eval mode

    @torch.no_grad()
    def query(self, feats: torch.Tensor, edges: np.ndarray, res: float, Lx: float, Ly: float, cx: float, cy: float):
        """cost_query_server.py:160-161 + CostQuery.__call__ (cost_query.py:39-69) + network.FCpart.
        edges [n,6] = [tx,ty,tyaw,sx,sy,syaw]; returns [n,3] = (power, time, 1-prob)."""
        p = self.p
        t = torch.as_tensor(np.asarray(edges, dtype=np.float64))
        t = t.clone()
        t[:, 0] -= cx; t[:, 1] -= cy; t[:, 3] -= cx; t[:, 4] -= cy
        t[:, :3] = t[:, :3] - t[:, 3:]
        feat_res = res * DOWNSAMPLE
        row_bias = int((Lx / res - 2 * MAP_CLIP) / DOWNSAMPLE * 0.5)
        col_bias = int((Ly / res - 2 * MAP_CLIP) / DOWNSAMPLE * 0.5)
        Hf, Wf = feats.shape[1], feats.shape[2]
        row = torch.clamp(t[:, 3] / feat_res + row_bias, min=1, max=Hf - 2).long()
        col = torch.clamp(t[:, 4] / feat_res + col_bias, min=1, max=Wf - 2).long()
        f = feats[:, row, col].t().contiguous()                       # [n, 48]
        tar = torch.cat((t[:, :3], t[:, 5:6]), dim=1).to(torch.float32)   # [dx, dy, dyaw, syaw]
        ang = tar[:, 2]
        ang = torch.where(ang > math.pi, ang - 2 * math.pi, ang)
        ang = torch.where(ang < -math.pi, ang + 2 * math.pi, ang)
        info = torch.stack((tar[:, 0], tar[:, 1], torch.sqrt(tar[:, 0] ** 2 + tar[:, 1] ** 2),
                            torch.atan2(tar[:, 1], tar[:, 0]), ang, torch.cos(ang), torch.sin(ang),
                            tar[:, 3], torch.cos(tar[:, 3]), torch.sin(tar[:, 3])), dim=1)   # [n, 10]

        def lin_bn(x, conv, bn):
            y = x @ p[conv + ".weight"].reshape(p[conv + ".weight"].shape[0], -1).t()
            s = p[bn + ".weight"] / torch.sqrt(p[bn + ".running_var"] + BN_EPS)
            return (y - p[bn + ".running_mean"]) * s + p[bn + ".bias"]

        def lin_bias(x, conv):
            return x @ p[conv + ".weight"].reshape(1, -1).t() + p[conv + ".bias"]

        tarf = lin_bn(info, "tar0_conv1", "tar0_conv1_bn")
        h = F.leaky_relu(lin_bn(torch.cat((f, tarf), dim=1), "out0_conv1", "out0_conv1_bn"), 0.3)
        power = F.relu(lin_bias(F.leaky_relu(lin_bn(h, "out1_conv1", "out1_conv1_bn"), 0.3), "out2_conv1"))
        tm = F.relu(lin_bias(F.leaky_relu(lin_bn(h, "out1_conv2", "out1_conv2_bn"), 0.3), "out2_conv2"))
        prob = torch.sigmoid(lin_bias(F.leaky_relu(lin_bn(h, "out1_conv3", "out1_conv3_bn"), 0.3), "out2_conv3"))
        return torch.cat((power, tm, 1.0 - prob), dim=1).numpy()
