"""Stand-in for the `yacs` package (not installable offline): the reference only needs `yacs.config.CfgNode`."""
