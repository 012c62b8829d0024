function J = JacobianPhi3D(x,N_m,xl,xu,yl,yu,zl,zu,Indices)
%JACOBIANPHI3D  Drop-in for tools/JacobianPhi3D.m (same signature, same output
% [3 x 3 x N_m x N]): the Hessians of the N_m basis functions at the columns of x are
% evaluated by librbslam's k_jacobian_phi3d through the MEX gateway.
% Put this directory before the reference's tools/ on the MATLAB path.
J = rbslam_mex('jacobianphi3d', double(x), double(N_m), xl, xu, yl, yu, zl, zu, double(Indices));
end
