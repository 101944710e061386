"""B200-native (sm_100a) hot path of ubc-vision/joint-regressor-refinement: batched SMPL
forward/backward fused with the 17x6890 joint regressor, the refinement loss, its analytic
backward and the per-pose Adam update, behind the reference's own Python interfaces."""
from . import synthetic  # noqa: F401
from . import utils  # noqa: F401
from ._lib import JrrError, LIB_PATH  # noqa: F401
from .discriminator import Discriminator, Shape_Discriminator  # noqa: F401
from .native import NativeModel, flatten_critic_state_dict  # noqa: F401
from .refine import CriticTrainer, PoseRefiner, RegressorRefit, export_normalised_regressor, load_j_regressor, save_j_regressor, shard_range  # noqa: F401
from .smpl import SMPL, SMPLFunction, SMPLOutput  # noqa: F401
from .utils import evaluate, find_j_reg_mask, find_joints, move_pelvis, rot6d_to_rotmat, set_seed  # noqa: F401
from .optimize import RefinementLoop  # noqa: F401
from .data import data_set, write_precomputed  # noqa: F401
from . import data  # noqa: F401
from .mesh_renderer import Mesh_Renderer, SilhouetteFunction, load_obj_faces, render_mesh, silhouette_mse  # noqa: F401
