function [traj_max,traj_mean] = particleFilterLocalization(dynModel,measModel,odometry,y,...
    x0_nonLin,Q,R,N_P,dt,makePlots) %#ok<INUSL>
%PARTICLEFILTERLOCALIZATION  Drop-in for examples/mag-localization-mapping/particleFilterLocalization.m:1-2
% running on the GPU (librbslam).  Same positional arguments and outputs.  dynModel / measModel are the
% handles of a model made by
%   m = rbslam_model('denseMag3D', NN, LL);  m = rbslam_localization_map(m, foo, dVarft, sigma2);
% where foo, dVarft, sigma2 are what the reference's measModel closure captures
% (run_localization.m:259-270).  R is unused, as in the reference.
  if nargin < 10, makePlots = []; end
  desc = rbslam_resolve(dynModel, measModel);
  if ~isfield(desc, 'foo')
    error('rbslam:unsupportedModel', 'attach the fixed map with rbslam_localization_map first');
  end
  if ~isempty(makePlots)
    warning('rbslam:makePlots', 'particleFilterLocalization: the per-step plotting hook is not forwarded');
  end
  opts = rbslam_opts();
  if strcmp(opts.rng, 'compat')   % pre-draw rand / randn in the reference's order (:90-97)
    N_T = size(y,1); U = zeros(N_P, N_T); Z = zeros(6, N_P, N_T);
    for t = 2:N_T
      for i = 1:N_P
        U(i,t) = rand; Z(1:3,i,t) = randn(3,1); Z(4:6,i,t) = randn(3,1);
      end
    end
    opts.U = U; opts.Z = Z;
  end
  [traj_max,traj_mean] = rbslam_mex('localization', desc, odometry, y, x0_nonLin, Q, N_P, dt, ...
                                    desc.foo, desc.dVarft, desc.sigma2, opts);
end
