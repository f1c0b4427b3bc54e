from . import models, update, algo
from .models import Actor, Critic
from .algo import Algo, DDPG, TD3
from .update import temporal_difference, value_update, ddpg_update, td3_update
