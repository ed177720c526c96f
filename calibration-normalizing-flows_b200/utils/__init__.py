"""Drop-in for the hot-path pieces of the reference's ``utils`` package."""
