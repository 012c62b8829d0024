function [XNK,XLK,PK] = particleSmoother(dynModel,measModel,dynResNorm,odometry,y,...
    x0_nonLin,x0_lin,P0_lin,Q,R,N_P,N_K,dt,sparseFeatures,makePlots)
% PARTICLESMOOTHER - drop-in for src/particleSmoother.m:1-2 running on the GPU (librbslam).
  if nargin < 14 || isempty(sparseFeatures), sparseFeatures = false; end
  if nargin < 15, makePlots = []; end
  desc = rbslam_resolve(dynModel, measModel, dynResNorm);
  opts = rbslam_opts();
  % per sweep the gateway calls makePlots(xnk,xlk,k,XNK,XLK,PK) and prints the progress line,
  % as src/particleSmoother.m:359-365 does
  opts.makePlots = makePlots;
  if strcmp(opts.rng, 'compat')
    opts = rbslam_streams(opts, desc, N_P, size(y,1), N_K, true);
  end
  [XNK,XLK,PK] = rbslam_mex('smoother', desc, 0, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, ...
                            N_P, N_K, dt, opts);
end
