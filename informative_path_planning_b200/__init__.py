"""B200-native reward evaluation for informative path planning.

Drop-in for the reward-evaluation hot path of expeditionary-robotics/informative-path-planning:
  gpmodel_library : GPModel, OnlineGPModel           (add_data / predict_value / posterior_samples)
  aq_library      : info_gain, mean_UCB, hotspot_info_UCB, mves, exp_improvement, naive, naive_value,
                    sample_max_vals, global_maximization, general_target  (+ batched reward_paths)
  mcts_library    : cMCTS, Tree, BeliefTree, Node    (tree logic on the host, rewards on the GPU)
  paths_library   : Path_Generator, Dubins_Path_Generator (restated without the `dubins` C library)
  obstacles       : FreeWorld, BlockWorld, BugTrap, ChannelWorld (in_obstacle + box lists for the native path generators)
  ipp_library     : the legacy monolith's GPModel / rewards / get_reward seam
  envmodel_library: Environment (Gaussian world synthesis + sample_value)
  robot_library   : Robot (mission loop: predict_max, myopic choice in one device call or cMCTS, observe, append)
  evaluation_library: Evaluation (per-step metrics; regret sweeps over all candidate paths as batched device calls)
  parallel        : query / path sharding across the GPUs of one box (torch.distributed, NCCL)

All numerics run in libipp_b200.so (hand-written CUDA for sm_100a, csrc/) behind the C ABI in
include/ipp_b200.h.  Importing the package without the built library raises ImportError: there is
no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when csrc/libipp_b200.so is missing)
from . import aq_library, gpmodel_library  # noqa: F401
from .gpmodel_library import GPModel, OnlineGPModel  # noqa: F401

__all__ = ['aq_library', 'gpmodel_library', 'GPModel', 'OnlineGPModel']
