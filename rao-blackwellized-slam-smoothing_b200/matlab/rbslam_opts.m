function opts = rbslam_opts()
%RBSLAM_OPTS  Optional environment switches; no new required arguments anywhere.
  opts.device = str2double(getenv_default('RBSLAM_DEVICE', '0'));
  opts.rng = getenv_default('RBSLAM_RNG', 'compat');
  opts.seed = str2double(getenv_default('RBSLAM_SEED', '0'));
  % RBSLAM_DEVICES='0,1,2,3': one filter sharded over these GPUs from this MATLAB process
  % (rbslam_create_group; particleFilter with the device RNG, setenv('RBSLAM_RNG','philox'))
  dv = getenv('RBSLAM_DEVICES');
  if ~isempty(dv), opts.devices = str2double(strsplit(dv, ',')); else, opts.devices = []; end
  opts.makePlots = [];
end
function v = getenv_default(name, dflt)
  v = getenv(name); if isempty(v), v = dflt; end
end
