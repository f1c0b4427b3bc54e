from .misc import temporal_difference, value_update
from .ddpg import ddpg_update
from .td3 import td3_update
from .reinforce import ChooseREINFORCE, reinforce_update
