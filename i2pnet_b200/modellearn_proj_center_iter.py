"""Iterative-refinement inference model (mirror of the reference's src/modellearn_proj_center_iter.py:
RegNet_v2 :24-440): the same network and state_dict as modellearn_proj_center.RegNet_v2, but the level-3
refinement (warp -> cost_volume2 -> flow predictors -> pose head -> compose, :345-404) runs six times.
Iteration 0 starts from the coarse pose (q4, t4); iteration i > 0 starts from the RESIDUAL pose (q3, t3)
regressed by iteration i - 1, exactly as the reference feeds `decalib_quat_real3 / dual3` back (:351-352);
the output is the composition computed in the last iteration.  Every iteration reuses the level-3 / level-4
features and the two up-convolutions; per iteration the cost is one kNN (228 x 80, k = 32), one cost volume,
two flow predictors and one pose head."""
from .config_proj_lidarcenter import I2PNetConfig as cfg_default
from .modellearn_proj_center import RegNet_v2 as _RegNet_v2
from .modellearn_proj_center import change_intrinsic, get_num_parameters, set_id_grid  # noqa: F401  (reference exports)

ITERATIONS = 6   # src/modellearn_proj_center_iter.py:346


class RegNet_v2(_RegNet_v2):
    def forward(self, rgb_img, lidar_img, lidar_img_raw, H_initial, intrinsic, resize_img, gt_project=None,
                calib=None, lidar_feature=None, cfg=cfg_default):
        s = self._coarse(rgb_img, lidar_img, lidar_img_raw, intrinsic, lidar_feature, cfg)
        q_in, t_in = s["q4"], s["t4"]
        for _ in range(ITERATIONS):
            out_3, W_l3, q_in, t_in = self._refine(s, q_in, t_in, cfg)
        return self._result(s, out_3, W_l3)
