def dice_loss(*a, **k): raise NotImplementedError
