registry = {}


def register(id, entry_point=None, **kwargs):
    registry[id] = {"id": id, "entry_point": entry_point, "kwargs": kwargs.get("kwargs", {})}
