"""Testsolver — the reference ships an empty stub (solver/testsolver.py:9-15) that reads a
non-existent cfg['algorithm'].  Here it is the evaluation-only solver: load weights, test, colour."""
from solver.mainsolver import Solver


class Testsolver(Solver):
    def __init__(self, cfg):
        cfg = dict(cfg)
        cfg['train'] = dict(cfg['train'], index=0)
        cfg.setdefault('model_name', cfg.get('algorithm', 'gmfnet'))
        super().__init__(cfg)
