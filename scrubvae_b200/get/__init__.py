from .model import model
