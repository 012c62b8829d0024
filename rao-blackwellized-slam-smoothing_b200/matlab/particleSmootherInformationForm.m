function [XNK,XLK,PK] = particleSmootherInformationForm(dynModel,measModel,dynResNorm,odometry,y,...
    x0_nonLin,x0_lin,P0_lin,Q,R,N_P,N_K,dt,sparseFeatures,makePlots)
% PARTICLESMOOTHERINFORMATIONFORM - drop-in for src/particleSmootherInformationForm.m:1-2.
  if nargin < 14 || isempty(sparseFeatures), sparseFeatures = false; end
  if nargin < 15, makePlots = []; end
  if sparseFeatures == true
    disp('This code has only been implemented for dense features')   % :77-80
    return;
  end
  desc = rbslam_resolve(dynModel, measModel, dynResNorm);
  opts = rbslam_opts();
  % per sweep the gateway calls makePlots(xnk,xlk,k,XNK,XLK,PK) and prints the progress line,
  % as src/particleSmoother.m:359-365 does
  opts.makePlots = makePlots;
  if strcmp(opts.rng, 'compat')
    opts = rbslam_streams(opts, desc, N_P, size(y,1), N_K, true);
  end
  [XNK,XLK,PK] = rbslam_mex('smoother', desc, 1, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, ...
                            N_P, N_K, dt, opts);
end
