def focal_loss(*a, **k): raise NotImplementedError
