from typing import Any

OptimizerLRScheduler = Any
