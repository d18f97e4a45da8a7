"""Runtime utilities (reference ``distribuuuu/utils.py`` surface, split by concern)."""
from .checkpoint import (count_parameters, get_checkpoint, get_checkpoint_dir, get_last_checkpoint,
                         has_checkpoint, load_checkpoint, save_checkpoint, unwrap_model)
from .data import (IMAGENET_MEAN, IMAGENET_STD, DummyDataset, PinnedPrefetcher, SyntheticDeviceLoader, construct_train_loader,
                   construct_val_loader, normalize_uint8)
from .dist import (barrier, broadcast_object, get_rank, get_world_size, is_primary, resolve_backend,
                   resolve_device, scaled_all_reduce, setup_distributed, shutdown)
from .env import setup_logger, setup_seed
from .health import Heartbeat, StepWatchdog, read_heartbeats, stale_ranks
from .lr import get_epoch_lr, get_lr_fun, lr_fun_cos, lr_fun_steps, set_lr
from .meters import AverageMeter, DeviceMetrics, ProgressMeter, accuracy, construct_meters
from .optim import construct_optimizer, sgd_hparams
from .timers import StepTimer
