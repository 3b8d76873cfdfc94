# -*- coding: utf8 -*-
"""Drop-in ``Config`` of the reference's ``offline/noahmp_config.py`` (same attribute names, same
``ERR:`` message + ``sys.exit(1)`` behaviour, reference lines cited), plus the mapping from namelist
option names to the 12 physics options the reference never wires up.

Differences, all additive: falls back to the in-repo namelist reader when the third-party ``f90nml`` is
not installed; keeps both the attribute pair the reference's ``__init__`` creates (``datetimebeg/end``,
cfg:59-60) and the pair its ``parse_cfg`` actually fills (``begdatetime/enddatetime``, cfg:99-101).
"""
import datetime
import os
import os.path
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
try:  # the reference's dependency, if present
    import f90nml
except ImportError:  # pragma: no cover - this image
    from noahmp_b200.compat import f90nml

# mandatory namelist entries, cfg:8-44
_PATHS = ['static_parameter_file', 'initialization_file', 'restart_file']
_IO = ['input_directory', 'input_frequency', 'output_directory', 'output_frequency', 'restart_directory',
       'restart_frequency']
_WHEN = [f'{edge}_{unit}' for edge in ('start', 'end') for unit in ('year', 'month', 'day', 'hour', 'minute', 'second')]
_PHYS = ['opt_veg', 'opt_run', 'opt_btr', 'opt_rad', 'opt_tub', 'opt_can', 'opt_inf', 'opt_snf', 'opt_tbot']
NML_FIELDS = _PATHS + _IO + _WHEN + ['interval_seconds'] + _PHYS

# namelist key -> physics option (global.f90:17-74); opt_tub is opt_sfc, opt_can is opt_crs
_OPTION_SOURCE = {'opt_veg': 'opt_veg', 'opt_crs': 'opt_can', 'opt_btr': 'opt_btr', 'opt_run': 'opt_run',
                  'opt_sfc': 'opt_tub', 'opt_inf': 'opt_inf', 'opt_rad': 'opt_rad', 'opt_snf': 'opt_snf',
                  'opt_tbot': 'opt_tbot'}
# not in the namelist: defaults from the comments at global.f90:45,60,73
_OPTION_DEFAULT = {'opt_frz': 1, 'opt_alb': 2, 'opt_stc': 1}
OPTION_ORDER = ['opt_veg', 'opt_crs', 'opt_btr', 'opt_run', 'opt_sfc', 'opt_frz', 'opt_inf', 'opt_rad', 'opt_alb',
                'opt_snf', 'opt_tbot', 'opt_stc']  # argument order of noahmp_set_options, global.f90:77-80


def _die(*msg):
    print('ERR:', *msg)
    sys.exit(1)


class Config(object):
    def __init__(self, cfgfile):
        # defaults of cfg:49-61
        self.indir, self.infreq = '.', None
        self.outdir, self.outfreq = '.', None
        self.resdir, self.resfreq = '.', None
        self.constfile, self.initfile = 'domain.nc', 'init.nc'
        self.datetimebeg = self.datetimeend = None
        self.timestep = 0
        self.parse_cfg(cfgfile)

    def parse_cfg(self, cfgfile):
        if not os.path.isfile(cfgfile):                                   # cfg:66-68
            _die('Unable to find configuration file', cfgfile)
        cfg = f90nml.read(cfgfile)['NOAHMP_OFFLINE']                      # cfg:69-70
        for var in NML_FIELDS:                                            # cfg:72-75
            if var not in cfg:
                _die('Unable to find {:s} in configuration file {:s}'.format(var, cfgfile))
        self.constfile, self.initfile, self.resfile = (cfg[k] for k in _PATHS)         # cfg:77-79
        (self.indir, self.infreq, self.outdir, self.outfreq,
         self.resdir, self.resfreq) = (cfg[k] for k in _IO)                            # cfg:81-86
        for key in _PHYS:                                                              # cfg:88-96
            setattr(self, key, cfg[key])
        self.timestep = datetime.timedelta(seconds=cfg['interval_seconds'])            # cfg:98
        self.begdatetime = datetime.datetime(*(cfg[k] for k in _WHEN[:6]))             # cfg:99-100
        self.enddatetime = datetime.datetime(*(cfg[k] for k in _WHEN[6:]))             # cfg:101-102
        self.datetimebeg, self.datetimeend = self.begdatetime, self.enddatetime
        # optional additions (absent keys keep the documented defaults)
        for key, default in _OPTION_DEFAULT.items():
            setattr(self, key, cfg.get(key, default))
        self.table_directory = cfg.get('table_directory', None)
        self.soil_table_tag = cfg.get('soil_table_tag', 'STAS')
        self.veg_table_tag = cfg.get('veg_table_tag', 'USGS')

    def physics_options(self):
        """The 12 integers of ``noahmp_set_options`` in the reference's argument order."""
        out = []
        for name in OPTION_ORDER:
            out.append(getattr(self, _OPTION_SOURCE[name]) if name in _OPTION_SOURCE else getattr(self, name))
        return out
