OptState = object  # only a type annotation in core/types.py
