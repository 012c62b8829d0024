function [XNK,XLK,PK] = particleSmoother(dynModel,measModel,dynResNorm,odometry,y,...
    x0_nonLin,x0_lin,P0_lin,Q,R,N_P,N_K,dt,sparseFeatures,makePlots)
% PARTICLESMOOTHER - drop-in for src/particleSmoother.m:1-2 running on the GPU (librbslam).
  if nargin < 14 || isempty(sparseFeatures), sparseFeatures = false; end
  if nargin < 15, makePlots = []; end
  desc = rbslam_resolve(dynModel, measModel, dynResNorm);
  opts = rbslam_opts();
  if strcmp(opts.rng, 'compat')
    opts = rbslam_streams(opts, desc, N_P, size(y,1), N_K, true);
  end
  [XNK,XLK,PK] = rbslam_mex('smoother', desc, 0, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, ...
                            N_P, N_K, dt, opts);
  if ~isempty(makePlots)
    for k = 1:N_K, makePlots(XNK(:,:,k), XLK(:,k), k, XNK, XLK, PK); end   % :360-362
  end
end
