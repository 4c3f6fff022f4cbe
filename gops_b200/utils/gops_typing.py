from typing import Any, Dict

import torch

DataDict = Dict[str, torch.Tensor]
InfoDict = Dict[str, Any]
ConfigDict = Dict[str, Any]
