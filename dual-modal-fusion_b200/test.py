"""Entry point, as the reference's test.py:7-14."""
import torch

from solver.mainsolver import Solver
from utils.config import get_render_config

if __name__ == "__main__":
    torch.manual_seed(3407)
    cfg = get_render_config("config.yml")
    Solver(cfg).run()
