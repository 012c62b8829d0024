function [traj_max,traj_mean,xl_max,xl_mean,P_max,P_mean,traj_sample_iwmax,xn_traj] = ...
    particleFilter(dynModel,measModel,odometry,y,...
    x0_nonLin,x0_lin,P0_lin,Q,R,N_P,dt,sparseFeatures,makePlots)
% PARTICLEFILTER - drop-in for src/particleFilter.m:1-3 running on the GPU (librbslam).
%
% Same positional arguments and outputs as the reference.  dynModel/measModel must be
% handles made by rbslam_model (the descriptor of the reference's closure family).
% Randomness: compat mode by default (rand/randn are pre-drawn from MATLAB's global
% stream in the reference's order); setenv('RBSLAM_RNG','philox') uses the device RNG.
  if nargin < 12 || isempty(sparseFeatures), sparseFeatures = false; end
  if nargin < 13, makePlots = []; end
  desc = rbslam_resolve(dynModel, measModel);
  if logical(sparseFeatures) ~= strcmp(desc.family, 'sparseVisual2D')
    error('rbslam:unsupportedModel', 'sparseFeatures does not match the model family');
  end
  opts = rbslam_opts();
  % makePlots(xn,xl(:,iw_max),P(:,:,iw_max),traj_max,yhattraj,xn_traj,traj_mean,xl,P) is called by the
  % gateway after every time step, as src/particleFilter.m:215-217 does
  opts.makePlots = makePlots;
  if strcmp(opts.rng, 'compat')
    opts = rbslam_streams(opts, desc, N_P, size(y,1), 1, false);
  end
  [traj_max,traj_mean,xl_max,xl_mean,P_max,P_mean,traj_sample_iwmax,xn_traj] = ...
      rbslam_mex('filter', desc, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P, dt, opts);
end
