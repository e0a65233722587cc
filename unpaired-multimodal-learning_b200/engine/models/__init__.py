from .head import UML, UMLClip, get_zero_shot_weights  # noqa: F401
