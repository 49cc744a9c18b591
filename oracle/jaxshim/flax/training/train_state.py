class TrainState:  # only a type annotation in core/types.py
    pass
