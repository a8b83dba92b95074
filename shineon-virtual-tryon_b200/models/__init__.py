"""Model lookup, mirroring models/__init__.py:4-32 of the reference."""
import importlib


def find_model_using_name(model_name):
    modellib = importlib.import_module(__name__ + "." + model_name + "_model")
    target = model_name.replace("_", "") + "model"
    from .base_model import BaseModel

    for name, cls in modellib.__dict__.items():
        if name.lower() == target.lower() and isinstance(cls, type) and issubclass(cls, BaseModel):
            return cls
    raise NotImplementedError(f"In {model_name}_model.py there should be a BaseModel subclass named {target}")


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options
