"""Host-side mirror of the reference's ``data`` package — only the step that sits directly in front of the training path."""
