"""GPU generators with the reference's generator protocol (supervillain/generator/generator.py:3-33)."""
from .generator import Generator
from . import villain, worldline, combining

__all__ = ['Generator', 'villain', 'worldline', 'combining']
