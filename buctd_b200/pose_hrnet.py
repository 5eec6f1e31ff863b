"""Drop-in for the reference's ``models.pose_hrnet`` (MODEL.NAME = 'pose_hrnet', plain HRNet or the
BUCTD preNet variant): ``get_pose_net(cfg, is_train, **kwargs)``,
/root/reference/lib/models/pose_hrnet.py:615-621."""
from .module import PoseHighResolutionNet


def get_pose_net(cfg, is_train, **kwargs):
    model = PoseHighResolutionNet(cfg, **kwargs)
    if is_train and cfg["MODEL"]["INIT_WEIGHTS"]:
        model.init_weights(cfg["MODEL"]["PRETRAINED"])
    return model
