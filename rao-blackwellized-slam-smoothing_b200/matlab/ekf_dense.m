function [xf_traj,qnb_traj,Pf_traj] = ekf_dense(dynModel,measModel,odometry,y,x0,q0,P0,Q,R,dt)
%EKF_DENSE  Drop-in for examples/slam-dense-mag/ekf_dense.m:1-2 running on the GPU (librbslam).
%
% Same positional arguments and outputs as the reference.  dynModel / measModel must be the
% handles m.dynModel_ekf / m.measModel_ekf of a model made by
%   m = rbslam_model('denseMag3D', NN, LL)        % LL = the 2 x 3 domain bounds
% (the closures dynModel_ekf / measModel_ekf of run_dense3D_magfield.m:281-316 are evaluated on
% the device; measModel_ekf's call of JacobianPhi3D uses LL as the reference does, :292-294).
  desc = rbslam_resolve(dynModel, measModel);
  if ~strcmp(desc.family, 'denseMag3D')
    error('rbslam:unsupportedModel', 'ekf_dense is defined for the dense magnetic-field model');
  end
  if isfield(desc, 'LL'), LL = desc.LL; else, LL = [-desc.L; desc.L]; end
  if nargout > 2
    [xf_traj,qnb_traj,Pf_traj] = rbslam_mex('ekf', desc, odometry, y, x0, q0, P0, Q, R, dt, LL);
  else
    [xf_traj,qnb_traj] = rbslam_mex('ekf', desc, odometry, y, x0, q0, P0, Q, R, dt, LL);
  end
end
