class Digraph:  # imported by core/evaluators/mcts/state.py for a debug dump that the path never calls
    def __init__(self, *a, **k):
        raise NotImplementedError("graphviz is not available")
