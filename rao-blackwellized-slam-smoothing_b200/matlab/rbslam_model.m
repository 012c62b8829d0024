function m = rbslam_model(family, varargin)
%RBSLAM_MODEL  Model descriptor whose handles the rbslam drop-ins recognise.
%
%   m = rbslam_model('denseMag3D',   NN, L)        % run_dense3D_magfield.m closures
%   m = rbslam_model('denseRadio2D', NN, L)        % run_dense2D_withHeading.m closures
%   m = rbslam_model('sparseVisual2D', nLandmarks, f, fp, fw)   % pfslam.m / measurement.m
%
% NN and L are what tools/domain_cartesian_dx.m returns / uses (index tuples and
% domain half-widths).  m.dynModel, m.measModel and m.dynResNorm are function
% handles that carry the descriptor in their workspace; pass them to
% particleFilter / particleSmoother exactly where the reference passes its own
% closures.  They cannot be evaluated on the host: the models run on the GPU.
  desc.family = family;
  switch family
    case {'denseMag3D','denseRadio2D'}
      desc.NN = double(varargin{1});
      L = varargin{2};
      if size(L,1) > 1   % the 2 x d domain bounds LL
        desc.LL = double(L);                                     % kept for measModel_ekf's JacobianPhi3D call
        L = (max(L,[],1) - min(L,[],1))/2;                       % domain_cartesian_dx.m:27-29
      end
      desc.L = double(L(:)');
    case 'sparseVisual2D'
      desc.nLandmarks = varargin{1};
      desc.camera = [varargin{2}, varargin{3}, varargin{4}];
    otherwise
      error('rbslam:unsupportedModel', 'unknown model family %s', family);
  end
  m = desc;
  m.dynModel   = @(varargin) rbslam_handle_stub(desc, 'dynModel');
  m.measModel  = @(varargin) rbslam_handle_stub(desc, 'measModel');
  if strcmp(family, 'denseMag3D')   % the EKF baseline's closures (run_dense3D_magfield.m:281-316)
    m.dynModel_ekf  = @(varargin) rbslam_handle_stub(desc, 'dynModel_ekf');
    m.measModel_ekf = @(varargin) rbslam_handle_stub(desc, 'measModel_ekf');
  end
  if strcmp(family, 'sparseVisual2D')
    m.dynResNorm = [];                                   % psslam.m:118 passes []
  else
    m.dynResNorm = @(varargin) rbslam_handle_stub(desc, 'dynResNorm');
  end
end

function rbslam_handle_stub(desc, role) %#ok<INUSD>
  error('rbslam:unsupportedModel', ...
        'rbslam model handles are evaluated on the GPU and cannot be called on the host');
end
