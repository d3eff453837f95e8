"""Empty stand-in so the reference modules' top-level `import matplotlib.pyplot`
succeeds in this container (matplotlib is not installed). The hot path never
calls into it (SURVEY.md §8c shim 1). Test infrastructure only."""
