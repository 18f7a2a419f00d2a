"""Mirror of the reference's `src` package for the accelerated path (models, losses, sliding-window inference)."""
