"""Drop-in for the reference's ``easy_boston_data`` module."""
from fastoptsolver_b200.datagen import generate_correlated_boston_like_data  # noqa: F401
