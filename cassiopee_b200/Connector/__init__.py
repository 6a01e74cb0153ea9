"""Connector module (chimera path only).  `import cassiopee_b200.Connector.PyTree as X`."""
