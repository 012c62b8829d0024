function opts = rbslam_opts()
%RBSLAM_OPTS  Optional environment switches; no new required arguments anywhere.
  opts.device = str2double(getenv_default('RBSLAM_DEVICE', '0'));
  opts.rng = getenv_default('RBSLAM_RNG', 'compat');
  opts.seed = str2double(getenv_default('RBSLAM_SEED', '0'));
end
function v = getenv_default(name, dflt)
  v = getenv(name); if isempty(v), v = dflt; end
end
