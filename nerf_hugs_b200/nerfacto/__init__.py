"""The torch twins' call surface (`/root/reference/nerfacto/{train,eval}.py`) over the C ABI of libhugs_b200.so.

    from nerf_hugs_b200.nerfacto.models import model_config_dict, model_dict, criterion_dict
    from nerf_hugs_b200.nerfacto.utils.checkpoint_utils import load_snapshot, save_snapshot

replace `from models import ...` / `from utils.checkpoint_utils import ...` of the reference scripts.
"""
