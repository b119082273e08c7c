def collect_env():
    return {}
def set_multi_processing(*a, **k):
    pass
