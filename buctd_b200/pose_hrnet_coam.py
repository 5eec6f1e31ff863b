"""Drop-in for the reference's ``models.pose_hrnet_coam`` (MODEL.NAME = 'pose_hrnet_coam'):
``get_pose_net(cfg, is_train, **kwargs)``, /root/reference/lib/models/pose_hrnet_coam.py:612-618."""
from .module import PoseHighResolutionNet


def get_pose_net(cfg, is_train, **kwargs):
    model = PoseHighResolutionNet(cfg, **kwargs)
    if is_train and cfg["MODEL"]["INIT_WEIGHTS"]:
        model.init_weights(cfg["MODEL"]["PRETRAINED"])
    return model
