class MMLogger:
    @staticmethod
    def get_instance(*a, **k):
        return MMLogger()
    @staticmethod
    def get_current_instance():
        return MMLogger()
    def info(self, *a, **k): pass
    def warning(self, *a, **k): pass
