def build_optim_wrapper(*a, **k):
    raise NotImplementedError
